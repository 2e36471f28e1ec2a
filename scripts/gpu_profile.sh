#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench command (per-launch device time), (2) full captures of
# the tensor-core GEMMs and (3) of the beam-scan kernels (k_prune, k_score_nodes), mid-sentence frames.
mkdir -p gpurun_out
S=${1:-1024}
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
  python bench.py --profile --steps 1 --sentences $S > gpurun_out/prof_launch.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
echo "== full capture (tensor-core GEMMs)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 3 -f -o gpurun_out/prof_tc \
  python bench.py --profile --steps 1 --sentences $S > gpurun_out/prof_full.log 2>&1; echo "rc=$?"
echo "== full capture (beam scan)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_prune|k_score_nodes|k_tc_lse_merge|k_tc_gather_split" -s 40 -c 4 -f -o gpurun_out/prof_scan \
  python bench.py --profile --steps 1 --sentences $S > gpurun_out/prof_scan.log 2>&1; echo "rc=$?"
ls -la gpurun_out/

#!/bin/bash
# ncu launch list (per-launch device time) + one full capture of the tensor-core GEMM kernels.
mkdir -p gpurun_out
S=${1:-1024}
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/launches.csv \
  python bench.py --profile --steps 1 --sentences $S > gpurun_out/prof_launch.log 2>&1; echo "rc=$?"
python scripts/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launch_summary.txt
echo "== full capture (tensor-core GEMMs, mid-sentence frames)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 4 -f -o gpurun_out/prof_tc \
  python bench.py --profile --steps 1 --sentences $S > gpurun_out/prof_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/

#!/bin/bash
# r02 evidence: launch lists of every workload + full ncu captures of the tensor-core kernels
mkdir -p gpurun_out
for w in cfg2 cfg3 cfg4 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$w.csv python bench.py --profile --steps 1 --sentences 1024 --workload $w > gpurun_out/prof_launch_$w.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$w.csv > gpurun_out/launch_summary_$w.txt; head -14 gpurun_out/launch_summary_$w.txt
done
echo "== full capture cfg5 rs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_lse_rs -s 20 -c 4 -f -o gpurun_out/prof_rs_cfg5 python bench.py --profile --steps 1 --sentences 1024 --workload cfg5 > gpurun_out/prof_rs_cfg5.log 2>&1; echo "rc=$?"
echo "== full capture cfg3"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 60 -c 8 -f -o gpurun_out/prof_tc_cfg3 python bench.py --profile --steps 1 --sentences 1024 --workload cfg3 > gpurun_out/prof_tc_cfg3.log 2>&1; echo "rc=$?"
echo "== full capture cfg2"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 3 -f -o gpurun_out/prof_tc_cfg2 python bench.py --profile --steps 1 --sentences 1024 --workload cfg2 > gpurun_out/prof_tc_cfg2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep

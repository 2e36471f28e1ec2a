#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (tc)"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -k "tc or odd or lockstep or quantized" > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -4
for w in cfg3 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${w}b.csv python bench.py --profile --steps 1 --sentences 1024 --workload $w > gpurun_out/prof_launch_${w}b.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_${w}b.csv > gpurun_out/launch_summary_${w}b.txt; head -6 gpurun_out/launch_summary_${w}b.txt
done

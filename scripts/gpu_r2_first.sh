#!/bin/bash
# Round-2 first GPU call: full -m gpu suite (new cfg5 / unsorted fixtures), TC error probe, baseline bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/test_gpu.log
echo "== probe"; timeout 600 python scripts/tc_error_probe.py 128 > gpurun_out/tc_error_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/tc_error_probe.txt
for w in cfg2 cfg3 cfg5 cfg4; do
  echo "== bench $w"; timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --cpu-baseline-sentences 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "rc=$?"; tail -c 1500 gpurun_out/bench_$w.json
done

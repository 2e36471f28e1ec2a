#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (tc)"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s -k "tc or odd or lockstep or guard" > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -8
for w in cfg3 cfg5; do
echo "== bench $w"; timeout 900 python bench.py --steps 5 --warmup 3 --workload $w --extra none --cpu-baseline-sentences 1 > gpurun_out/bench_${w}_rs3.json 2> gpurun_out/bench_${w}_rs3.err; echo "rc=$?"; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_${w}_rs3.json') if l.startswith('{')][-1]; r=d['roofline']; print(d['value'], d['ms_per_step'], 'off', d['guard']['value_guard_off'], 'roof', r['frac'], r['avg_launch_ms'], r['launches'], d['cpu_baseline']['nbest_identical_to_gpu'], d['clocks'])"; tail -2 gpurun_out/bench_${w}_rs3.err
done

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (tc)"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s -k "tc or odd or guard" > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error" gpurun_out/test_gpu.log | tail -20
for w in cfg3 cfg5; do
echo "== bench $w rs"; timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --extra none --cpu-baseline-sentences 1 > gpurun_out/bench_${w}_rs.json 2> gpurun_out/bench_${w}_rs.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_rs.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'roof', r['frac'], r['avg_launch_ms'], r['launches'], 'guard', {k:v for k,v in d['guard'].items() if k!='note'}, d['cpu_baseline']['nbest_identical_to_gpu'])"; tail -2 gpurun_out/bench_${w}_rs.err
echo "== bench $w old"; JLM_TC_RS=0 JLM_GUARD_EPS=0 timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --extra none --cpu-baseline-sentences 1 > gpurun_out/bench_${w}_old.json 2> gpurun_out/bench_${w}_old.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_old.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'roof', r['frac'], r['avg_launch_ms'], r['launches'])"
done

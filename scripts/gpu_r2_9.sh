#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error" gpurun_out/test_gpu.log | tail -20
for w in cfg3 cfg5 cfg2; do
echo "== bench $w"; JLM_GUARD_EPS=0 timeout 900 python bench.py --steps 3 --warmup 3 --workload $w --extra none --cpu-baseline-sentences 1 > gpurun_out/bench_${w}_rs2.json 2> gpurun_out/bench_${w}_rs2.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_rs2.json')); r=d['roofline']; print(d['value'], d['ms_per_step'], 'roof', r['frac'], r['avg_launch_ms'], r['launches'], 'lat', d['config']['single_sentence_latency_ms'], d['cpu_baseline']['nbest_identical_to_gpu'])"; tail -2 gpurun_out/bench_${w}_rs2.err
done
echo "== latency old skinny"; JLM_STREAM_GEMM=0 JLM_GUARD_EPS=0 timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg2 --extra none --cpu-baseline-sentences 1 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lat', d['config']['single_sentence_latency_ms'])"

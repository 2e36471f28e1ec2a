#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "guard|bound|passed|failed|Error|error" gpurun_out/test_gpu.log | tail -20
echo "== bench full"; JLM_DEBUG_TIMING=1 timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_new.json'))
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), 'guard', {k: v for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'])"; grep guard gpurun_out/bench_new.err | sort | uniq -c | sort -rn | head -5; grep -v guard gpurun_out/bench_new.err | tail -5
echo "== bench guard off"; JLM_GUARD_EPS=0 timeout 600 python bench.py --steps 10 --warmup 3 --extra none --cpu-baseline-sentences 2 > gpurun_out/bench_new_noguard.json 2> gpurun_out/bench_new_noguard.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_new_noguard.json')); print(d['value'], d['e2e']['value'], d['e2e']['blocking_value'], d['ms_per_step'], d['clocks'])"

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -14
echo "== bench full"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_new4.json 2> gpurun_out/bench_new4.err; echo "rc=$?"; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_new4.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), x['roofline'].get('avg_launch_ms'), 'gate', (x['roofline'].get('gate') or {}).get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'])"; tail -3 gpurun_out/bench_new4.err

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -14
echo "== bench full"; JLM_DEBUG_TIMING=1 timeout 1200 python bench.py --steps 10 --warmup 3 --e2e-depth ${DEPTH:-2} > gpurun_out/bench_new3.json 2> gpurun_out/bench_new3.err; echo "rc=$?"; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_new3.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'])"; grep guard gpurun_out/bench_new3.err | sort | uniq -c | sort -rn | head -6
echo "== e2e depth 3"; timeout 600 python bench.py --steps 10 --warmup 3 --extra none --cpu-baseline-sentences 2 --e2e-depth 3 2>/dev/null | python -c "
import json,sys; d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]; print('depth3', d['value'], d['e2e']['value'], d['e2e']['blocking_value'])"

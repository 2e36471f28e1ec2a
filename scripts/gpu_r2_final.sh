#!/bin/bash
# what the driver runs at round end, timed: reference arm, our arm, smoke
mkdir -p gpurun_out
/usr/bin/time -v python -c "import __graft_entry__ as g; g.smoke()" 2> gpurun_out/smoke.time | tail -2; grep -E "Elapsed|Maximum resident" gpurun_out/smoke.time
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json ) 2>&1 | grep real; cut -c1-400 gpurun_out/bench_ref.json
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), x['roofline'].get('avg_launch_ms'), 'gate', (x['roofline'].get('gate') or {}).get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'], 'launches', d['gpu_launches'])"; tail -3 gpurun_out/bench_final.err

#!/bin/bash
# last check of the round: GPU tests, our arm with the driver's arguments, launch lists of the workloads whose kernels changed last
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -7
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), x['roofline'].get('avg_launch_ms'), 'gate', (x['roofline'].get('gate') or {}).get('frac'), 'guard_off %.2f' % x['guard']['ms_per_step_guard_off'], 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'], 'launches', d['gpu_launches'])"; tail -3 gpurun_out/bench_final.err
for w in cfg2 cfg4 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$w.csv python bench.py --profile --steps 1 --sentences 1024 --workload $w > gpurun_out/prof_launch_$w.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$w.csv > gpurun_out/launch_summary_$w.txt; head -9 gpurun_out/launch_summary_$w.txt
done

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests (guard)"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s -k "guard or lockstep or stream or odd" > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -12
echo "== bench full"; JLM_DEBUG_TIMING=1 timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_new2.json 2> gpurun_out/bench_new2.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_new2.json'))
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'])"; grep guard gpurun_out/bench_new2.err | sort | uniq -c | sort -rn | head -6
echo "== single launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_single2.csv python scripts/latency_probe.py 2 > gpurun_out/prof_launch_single2.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_single2.csv | head -8
JLM_STREAM_GEMM=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_single3.csv python scripts/latency_probe.py 2 > gpurun_out/prof_launch_single3.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_single3.csv | head -8

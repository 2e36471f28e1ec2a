#!/bin/bash
# what the driver runs at round end, timed: smoke, reference arm, our arm; then the final evidence captures
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), x['roofline'].get('avg_launch_ms'), 'gate', (x['roofline'].get('gate') or {}).get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'], 'launches', d['gpu_launches'])"; tail -3 gpurun_out/bench_final.err
for w in cfg2 cfg3 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$w.csv python bench.py --profile --steps 1 --sentences 1024 --workload $w > gpurun_out/prof_launch_$w.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$w.csv > gpurun_out/launch_summary_$w.txt; head -8 gpurun_out/launch_summary_$w.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_lse_rs -s 20 -c 2 -f -o gpurun_out/prof_rs_cfg5_final python bench.py --profile --steps 1 --sentences 1024 --workload cfg5 > gpurun_out/prof_rs_cfg5.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_ -s 60 -c 8 -f -o gpurun_out/prof_tc_cfg3_final python bench.py --profile --steps 1 --sentences 1024 --workload cfg3 > gpurun_out/prof_tc_cfg3.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 3 -f -o gpurun_out/prof_tc_cfg2_final python bench.py --profile --steps 1 --sentences 1024 --workload cfg2 > gpurun_out/prof_tc_cfg2.log 2>&1; echo "rc=$?"

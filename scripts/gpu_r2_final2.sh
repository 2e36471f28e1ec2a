#!/bin/bash
# what the driver runs at round end, timed: GPU tests, smoke, reference arm, our arm; then the evidence captures of this session's kernels
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -7
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -4
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_final.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM (blocking %.3fM) ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['e2e']['blocking_value']/1e6, x['ms_per_step']), 'roof', x['roofline'].get('frac'), x['roofline'].get('avg_launch_ms'), 'gate', (x['roofline'].get('gate') or {}).get('frac'), 'guard', {k: (round(v,4) if isinstance(v,float) else v) for k, v in x['guard'].items() if k != 'note'}, 'cpu', x['cpu_baseline']['value'], x['cpu_baseline']['nbest_identical_to_gpu'], 'strong', (x.get('strong') or {}).get('value'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'], 'lat', d['config']['single_sentence_latency_ms'], 'launches', d['gpu_launches'], d['config']['single_sentence_roofline'])"; tail -3 gpurun_out/bench_final.err
for w in cfg2 cfg3 cfg4 cfg5; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$w.csv python bench.py --profile --steps 1 --sentences 1024 --workload $w > gpurun_out/prof_launch_$w.log 2>&1
  python scripts/summarize_launches.py gpurun_out/launches_$w.csv > gpurun_out/launch_summary_$w.txt; head -9 gpurun_out/launch_summary_$w.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_vocab_dense -s 20 -c 2 -f -o gpurun_out/prof_vocab_dense python bench.py --profile --steps 1 --sentences 1024 --workload cfg4 > gpurun_out/prof_vocab_dense.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_single_f64 -s 4 -c 1 -f -o gpurun_out/prof_single python scripts/latency_probe.py 4 0 > gpurun_out/prof_single.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 3 -f -o gpurun_out/prof_tc_cfg2_final python bench.py --profile --steps 1 --sentences 1024 --workload cfg2 > gpurun_out/prof_tc_cfg2.log 2>&1; echo "rc=$?"
python scripts/latency_probe.py 40 0 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "lock-step|guard|bound|passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -20
python scripts/latency_probe.py 20
JLM_STREAM_GEMM=0 python scripts/latency_probe.py 20
echo "== bench cfg4"; JLM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg4 --extra none --cpu-baseline-sentences 2 > gpurun_out/bench_cfg4_sh.json 2> gpurun_out/bench_cfg4_sh.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_sh.json')); r=d['roofline']; print(d['value'], d['e2e']['value'], d['ms_per_step'], 'guard', {k:v for k,v in d['guard'].items() if k!='note'}, d['cpu_baseline']['nbest_identical_to_gpu'], 'strong', d.get('strong',{}).get('value'))"; grep guard gpurun_out/bench_cfg4_sh.err | tail -2; grep -v guard gpurun_out/bench_cfg4_sh.err | tail -3
echo "== launch list cfg4"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --profile --steps 1 --sentences 1024 --workload cfg4 > gpurun_out/prof_launch_cfg4.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_cfg4.csv | head -16

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -8
echo "== bench cfg4"; timeout 900 python bench.py --steps 5 --warmup 3 --workload cfg4 --extra none --cpu-baseline-sentences 2 > gpurun_out/bench_cfg4_v3.json 2> gpurun_out/bench_cfg4_v3.err; echo "rc=$?"; python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_cfg4_v3.json') if l.startswith('{')][-1]; print(d['value'], d['e2e']['value'], d['ms_per_step'], 'off', d['guard']['value_guard_off'], d['cpu_baseline']['nbest_identical_to_gpu'], d['clocks'])"; tail -2 gpurun_out/bench_cfg4_v3.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_cfg4b.csv python bench.py --profile --steps 1 --sentences 1024 --workload cfg4 > gpurun_out/prof_launch_cfg4b.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_cfg4b.csv > gpurun_out/launch_summary_cfg4b.txt; head -10 gpurun_out/launch_summary_cfg4b.txt

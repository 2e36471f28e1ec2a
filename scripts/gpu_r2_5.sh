#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x -s > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "guard|bound|passed|failed|Error|error" gpurun_out/test_gpu.log | tail -20
echo "== parity at scale"; timeout 900 python scripts/parity_at_scale.py 1024 > gpurun_out/parity_at_scale2.txt 2>&1; echo "rc=$?"; cat gpurun_out/parity_at_scale2.txt
echo "== probe"; timeout 600 python scripts/tc_error_probe.py 128 > gpurun_out/tc_error_probe4.txt 2>&1; echo "rc=$?"; tail -6 gpurun_out/tc_error_probe4.txt
echo "== bench cfg2"; timeout 600 python bench.py --steps 10 --warmup 3 --cpu-baseline-sentences 2 > gpurun_out/bench_cfg2_c.json 2> gpurun_out/bench_cfg2_c.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_c.json')); print(d['value'], d['e2e']['value'], d['e2e']['blocking_value'], d['ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench_cfg2_c.err
echo "== bench cfg2 guard off"; JLM_GUARD_EPS=0 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-baseline-sentences 2 > gpurun_out/bench_cfg2_d.json 2> gpurun_out/bench_cfg2_d.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_d.json')); print(d['value'], d['e2e']['value'], d['e2e']['blocking_value'], d['ms_per_step'], d['clocks'])"

#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_gpu.log
echo "== probe"; timeout 600 python scripts/tc_error_probe.py 128 > gpurun_out/tc_error_probe2.txt 2>&1; echo "rc=$?"; cat gpurun_out/tc_error_probe2.txt
echo "== probe no debias"; JLM_TC_DEBIAS=0 timeout 600 python scripts/tc_error_probe.py 128 > gpurun_out/tc_error_probe2_nodebias.txt 2>&1; echo "rc=$?"; head -8 gpurun_out/tc_error_probe2_nodebias.txt; tail -12 gpurun_out/tc_error_probe2_nodebias.txt
echo "== parity at scale"; timeout 900 python scripts/parity_at_scale.py 1024 > gpurun_out/parity_at_scale.txt 2>&1; echo "rc=$?"; cat gpurun_out/parity_at_scale.txt
echo "== bench cfg2"; timeout 600 python bench.py --steps 5 --warmup 3 --cpu-baseline-sentences 2 > gpurun_out/bench_cfg2_b.json 2> gpurun_out/bench_cfg2_b.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_cfg2_b.json; tail -3 gpurun_out/bench_cfg2_b.err

#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 4 7; do
echo "== probe debias mask $m"; JLM_TC_DEBIAS=$m timeout 600 python scripts/tc_error_probe.py 128 > gpurun_out/tc_error_probe3_m$m.txt 2>&1; echo "rc=$?"; sed -n 1,3p gpurun_out/tc_error_probe3_m$m.txt; sed -n 12,14p gpurun_out/tc_error_probe3_m$m.txt; tail -5 gpurun_out/tc_error_probe3_m$m.txt
done
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_gpu.log

#!/bin/bash
mkdir -p gpurun_out
echo "== bias probe"; timeout 600 python scripts/tc_bias_probe.py > gpurun_out/tc_bias_probe.txt 2>&1; echo "rc=$?"; cat gpurun_out/tc_bias_probe.txt
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/test_gpu.log

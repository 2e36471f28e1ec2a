#!/bin/bash
# One gpurun call: exact-path parity, tensor-core parity, smoke, short bench. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== exact" ; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 \
  -k "exact or model_api or batch_equals or ragged" > gpurun_out/test_exact.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/test_exact.log
echo "== tc" ; timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 -s \
  -k "tc" > gpurun_out/test_tc.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/test_tc.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/bench.log

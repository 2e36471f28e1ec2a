#!/bin/bash
# why the guard's float64 re-scoring is slower under torchrun: same 1-GPU job, plain python vs torchrun, with / without OMP_NUM_THREADS=1
mkdir -p gpurun_out
echo "== plain"; JLM_DEBUG_TIMING=1 python bench.py --steps 5 --workload cfg4 --extra none --cpu-baseline-sentences 1 2>&1 | grep -E "guard:" | tail -3
echo "== plain OMP=1"; OMP_NUM_THREADS=1 JLM_DEBUG_TIMING=1 python bench.py --steps 5 --workload cfg4 --extra none --cpu-baseline-sentences 1 2>&1 | grep -E "guard:" | tail -3
echo "== torchrun 1"; JLM_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 1 --steps 5 --workload cfg4 --extra none --cpu-baseline-sentences 1 2>&1 | grep -E "guard:" | tail -3
nproc; python -c "import os; print(os.sched_getaffinity(0).__len__())"

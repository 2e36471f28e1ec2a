#!/bin/bash
# full ncu capture of the gate GEMM (k_tc_gemm<256,2>) and the projection GEMM (<256,0>), mid-sentence frames
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 3 -f -o gpurun_out/prof_tc \
  python bench.py --profile --steps 1 --sentences 1024 > gpurun_out/prof_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out/

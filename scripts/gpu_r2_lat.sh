#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_skinny_f64 -s 30 -c 3 -f -o gpurun_out/prof_skinny_hot python scripts/latency_probe.py 2 > gpurun_out/prof_skinny_hot.log 2>&1; echo "rc=$?"
JLM_STREAM_GEMM=2 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:k_stream_f64 -s 10 -c 2 -f -o gpurun_out/prof_stream_hot python scripts/latency_probe.py 2 > gpurun_out/prof_stream_hot.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*hot*

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_single_f64 -s 4 -c 1 -f -o gpurun_out/prof_single python scripts/latency_probe.py 4 0 > gpurun_out/prof_single.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/prof_single.log

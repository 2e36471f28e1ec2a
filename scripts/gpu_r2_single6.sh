#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "single_sentence or batch_equals_single or ragged or near_tie" 2>&1 | tail -3 )
timeout 120 python scripts/latency_probe.py 40 0 2>&1 | tail -4
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_single_f64 -s 4 -c 1 -f -o gpurun_out/prof_single python scripts/latency_probe.py 4 0 > gpurun_out/prof_single.log 2>&1; echo "rc=$?"

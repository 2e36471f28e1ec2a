#!/bin/bash
mkdir -p gpurun_out
python scripts/latency_probe.py 20
JLM_STREAM_GEMM=0 python scripts/latency_probe.py 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_f64 -s 40 -c 2 -f -o gpurun_out/prof_stream python scripts/latency_probe.py 2 > gpurun_out/prof_stream.log 2>&1; echo "rc=$?"
JLM_STREAM_GEMM=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_skinny_f64 -s 120 -c 2 -f -o gpurun_out/prof_skinny python scripts/latency_probe.py 2 > gpurun_out/prof_skinny.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_single.csv python scripts/latency_probe.py 2 > gpurun_out/prof_launch_single.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_single.csv
ls -la gpurun_out/*.ncu-rep

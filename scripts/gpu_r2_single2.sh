#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do echo "== JLM_SINGLE=$v"; JLM_SINGLE=$v timeout 120 python scripts/latency_probe.py 40 0 2>&1 | tail -5; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_single.csv python scripts/latency_probe.py 4 0 > gpurun_out/prof_single.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_single.csv | head -20

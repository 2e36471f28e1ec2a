#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 -x > gpurun_out/test_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed|Error|error|assert" gpurun_out/test_gpu.log | tail -8
python scripts/latency_probe.py 30
JLM_WARPCOL=0 python scripts/latency_probe.py 30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_single4.csv python scripts/latency_probe.py 2 > gpurun_out/prof_launch_single4.log 2>&1; python scripts/summarize_launches.py gpurun_out/launches_single4.csv 2>/dev/null | head -12

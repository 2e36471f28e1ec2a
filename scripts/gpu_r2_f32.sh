#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
P='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
def show(n, x): print(n, "value %.3fM e2e %.3fM (blk %.3fM) ms %.2f off %.2f" % (x["value"]/1e6, x["e2e"]["value"]/1e6, x["e2e"]["blocking_value"]/1e6, x["ms_per_step"], x["guard"]["ms_per_step_guard_off"]), "guard", x["guard"]["flagged_sentences_per_step"], x["guard"]["sentences_redecoded_f64_per_step"], "par", x["cpu_baseline"]["nbest_identical_to_gpu"])
show(d["config"]["workload"][:4], d)
print(d["clocks"])'
for rep in 1 2; do
  timeout 600 python bench.py --steps 10 --workload cfg4 --cpu-baseline-sentences 4 --extra none > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "== cfg4 rc=$?"; python -c "$P" < gpurun_out/bench_ab.json; tail -2 gpurun_out/bench_ab.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --profile --steps 1 --sentences 1024 --workload cfg4 > gpurun_out/prof_launch_cfg4.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_cfg4.csv > gpurun_out/launch_summary_cfg4.txt; head -10 gpurun_out/launch_summary_cfg4.txt

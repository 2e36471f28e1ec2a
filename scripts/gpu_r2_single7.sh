#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 )
timeout 120 python scripts/latency_probe.py 40 0 2>&1 | tail -4
P='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
def show(n, x): print(n, "value %.3fM e2e %.3fM (blk %.3fM) ms %.2f off %.2f" % (x["value"]/1e6, x["e2e"]["value"]/1e6, x["e2e"]["blocking_value"]/1e6, x["ms_per_step"], x["guard"]["ms_per_step_guard_off"]), "roof", round(x["roofline"].get("frac"),4), x["roofline"].get("avg_launch_ms"), "par", x["cpu_baseline"]["nbest_identical_to_gpu"])
show(d["config"]["workload"][:4], d)
for w in d["workloads"]: show(w["workload"], w)
print(d["clocks"], "lat", d["config"]["single_sentence_latency_ms"])'
for v in 1 0; do
  JLM_SINGLE=$v JLM_DEBUG_TIMING=1 timeout 600 python bench.py --steps 6 --workload cfg5 --cpu-baseline-sentences 1 --extra none > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "== JLM_SINGLE=$v rc=$?"; python -c "$P" < gpurun_out/bench_ab.json; grep "tier 2" gpurun_out/bench_ab.err | tail -2
done

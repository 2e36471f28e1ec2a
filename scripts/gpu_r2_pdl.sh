#!/bin/bash
# programmatic dependent launch: GPU parity tests with it on, then A/B of the bench (all workloads) with JLM_PDL=1 / 0
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
P='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
def show(n, x): print(n, "value %.3fM e2e %.3fM ms %.2f off %.2f" % (x["value"]/1e6, x["e2e"]["value"]/1e6, x["ms_per_step"], x["guard"]["ms_per_step_guard_off"]), "roof", round(x["roofline"].get("frac"),4), x["roofline"].get("avg_launch_ms"), "par", x["cpu_baseline"]["nbest_identical_to_gpu"])
show("cfg2", d)
for w in d["workloads"]: show(w["workload"], w)
print(d["clocks"], "lat", d["config"]["single_sentence_latency_ms"])'
for rep in 1 2; do
  for v in 1 0; do
    JLM_PDL=$v timeout 600 python bench.py --steps 10 --cpu-baseline-sentences 4 > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err; echo "== JLM_PDL=$v rc=$?"; python -c "$P" < gpurun_out/bench_pdl$v.json; tail -2 gpurun_out/bench_pdl$v.err
  done
done
JLM_DEBUG_TIMING=1 timeout 300 python bench.py --steps 3 --workload cfg5 --extra none --cpu-baseline-sentences 1 2>&1 | grep -E "guard|rerun" | sort | uniq -c | sort -rn | head -20

#!/bin/bash
# steady-state e2e (20 steps) of the launch-chaining variants
mkdir -p gpurun_out
P='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
def show(n, x): print(n, "value %.3fM e2e %.3fM (blk %.3fM) ms %.2f off %.2f" % (x["value"]/1e6, x["e2e"]["value"]/1e6, x["e2e"]["blocking_value"]/1e6, x["ms_per_step"], x["guard"]["ms_per_step_guard_off"]), "roof", round(x["roofline"].get("frac"),4), x["roofline"].get("avg_launch_ms"), "par", x["cpu_baseline"]["nbest_identical_to_gpu"])
show("cfg2", d)
for w in d["workloads"]: show(w["workload"], w)
print(d["clocks"], "lat", d["config"]["single_sentence_latency_ms"])'
for rep in 1 2; do
for v in "JLM_COPY_STREAM=0" "JLM_COPY_STREAM=1" "JLM_COPY_STREAM=0 JLM_PDL_DRAIN=0" "JLM_COPY_STREAM=0 JLM_PDL=0"; do
  env $v JLM_BENCH_E2E_DEBUG=1 timeout 600 python bench.py --steps 20 --cpu-baseline-sentences 4 --extra cfg5 --extra-steps 6 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "== $v rc=$?"; python -c "$P" < gpurun_out/bench_ab.json; grep "bench\] e2e" gpurun_out/bench_ab.err | grep -v "depth 1"
done
done

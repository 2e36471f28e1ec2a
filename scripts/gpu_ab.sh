#!/bin/bash
# A/B of one engine switch on the bench workload: usage gpu_ab.sh ENV_VAR  (runs VAR unset vs VAR=0, twice)
mkdir -p gpurun_out
V=${1:-JLM_TC_PAIR}
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), "proj_frac", round(r["frac"],4), "proj_ms", round(r["avg_launch_ms"],4), "gate_ms", round(r["gate"]["avg_launch_ms"],4), d["clocks"]["sm_mhz"], d["cpu_baseline"]["nbest_identical_to_gpu"])'
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 > gpurun_out/bench_on.log 2>&1; echo -n "on   "; tail -1 gpurun_out/bench_on.log | python -c "$P"
  env $V=0 timeout 300 python bench.py --steps 20 > gpurun_out/bench_off.log 2>&1; echo -n "off  "; tail -1 gpurun_out/bench_off.log | python -c "$P"
done

#!/bin/bash
# A/B of the CTA-pair (cta_group::2) GEMM against the single-CTA kernel on the bench workload
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), "proj_frac", round(r["frac"],4), "proj_ms", round(r["avg_launch_ms"],4), "gate_ms", round(r["gate"]["avg_launch_ms"],4), d["clocks"], d["cpu_baseline"]["nbest_identical_to_gpu"])'
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 > gpurun_out/bench_pair.log 2>&1; echo -n "pair   "; tail -1 gpurun_out/bench_pair.log | python -c "$P"
  JLM_TC_PAIR=0 timeout 300 python bench.py --steps 20 > gpurun_out/bench_nopair.log 2>&1; echo -n "single "; tail -1 gpurun_out/bench_nopair.log | python -c "$P"
done

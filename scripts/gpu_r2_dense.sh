#!/bin/bash
# dense TMA-fed vocabulary logits (cfg4) A/B, adaptive launch chaining (cfg2 / cfg5 e2e)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -7
P='import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith("{")][-1]
def show(n, x): print(n, "value %.3fM e2e %.3fM (blk %.3fM) ms %.2f off %.2f" % (x["value"]/1e6, x["e2e"]["value"]/1e6, x["e2e"]["blocking_value"]/1e6, x["ms_per_step"], x["guard"]["ms_per_step_guard_off"]), "roof", round(x["roofline"].get("frac"),4), x["roofline"].get("avg_launch_ms"), "par", x["cpu_baseline"]["nbest_identical_to_gpu"])
show(d["config"]["workload"][:4], d)
for w in d["workloads"]: show(w["workload"], w)
print(d["clocks"], "lat", d["config"]["single_sentence_latency_ms"])'
for rep in 1 2; do
for v in "JLM_TC_VOCAB_DENSE=1" "JLM_TC_VOCAB_DENSE=0"; do
  env $v timeout 600 python bench.py --steps 10 --workload cfg4 --cpu-baseline-sentences 4 --extra none > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "== $v rc=$?"; python -c "$P" < gpurun_out/bench_ab.json; tail -2 gpurun_out/bench_ab.err
done
done
for v in "JLM_X=1" "JLM_PDL=0"; do
  env $v timeout 600 python bench.py --steps 20 --cpu-baseline-sentences 4 --extra cfg3,cfg5 --extra-steps 6 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "== $v rc=$?"; python -c "$P" < gpurun_out/bench_ab.json; tail -2 gpurun_out/bench_ab.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --profile --steps 1 --sentences 1024 --workload cfg4 > gpurun_out/prof_launch_cfg4.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_cfg4.csv > gpurun_out/launch_summary_cfg4.txt; head -16 gpurun_out/launch_summary_cfg4.txt

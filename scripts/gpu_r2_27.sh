#!/bin/bash
mkdir -p gpurun_out
JLM_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --extra none --cpu-baseline-sentences 2 > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err; echo "rc=$?"
grep "contradicted" gpurun_out/bench_dbg.err | sort | uniq -c | head
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_dbg.json') if l.startswith('{')][-1]; print(d['value'], d['e2e']['value'], d['config']['single_sentence_latency_ms'], json.dumps(d['config']['single_sentence_roofline'])[:600])"
for seed in 1 2 3; do JLM_DEBUG_TIMING=1 python scripts/parity_at_scale.py 1024 $seed 2>&1 | grep -E "contradicted|eps 1.0e-04" | sort | uniq -c | head -6; done

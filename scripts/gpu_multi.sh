#!/bin/bash
# bench.py on N GPUs of one box (weak scaling line + strong-scaling arms), N = $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]
def show(n, x): print(n, 'value %.3fM e2e %.3fM ms %.2f' % (x['value']/1e6, x['e2e']['value']/1e6, x['ms_per_step']), 'guard', x['guard']['flagged_fraction'], x['guard']['sentences_redecoded_f64_per_step'], 'strong', (x.get('strong') or {}).get('value'), (x.get('strong') or {}).get('ms_per_step'))
show('cfg2', d)
for w in d['workloads']: show(w['workload'], w)
print(d['clocks'])"
tail -5 gpurun_out/bench_n$N.err

#!/usr/bin/env python
"""Single-sentence latency (exact float64 back end, cfg2 shape) with float32 output weights vs the
8-bit code + codebook form of train/comp.py.  Run on the GPU box: python scripts/latency_q8.py"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import config, synth  # noqa: E402

root = tempfile.mkdtemp(prefix='jlm_q8_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
dump, decoded = synth.write_compressed(root, 1, {k: v for k, v in weights.items()}, bits=8)
synth.write_experiment(root, 2, cfg, decoded)
config.set_root(root)
sents = synth.make_sentences(lexicon, 8, min_len=20, seed=5, vocab_size=50000)
os.environ['JLM_Q8'] = '1'      # read by jlm_create: stream the codes even though this block is L2-resident
for name, dec in (('float32 blocks', jlm_b200.Decoder(2)), ('8-bit codes', jlm_b200.Decoder(1, comp=8))):
    out = [dec.decode_batch([s], backend=1) for s in sents]          # warm-up
    t0 = time.perf_counter()
    for _ in range(5):
        res = [dec.decode_batch([s], backend=1) for s in sents]
    dt = (time.perf_counter() - t0) / (5 * len(sents))
    print('%-16s %.3f ms / sentence (%d kana)  quantized=%s  top1=%s' % (name, dt * 1e3, len(sents[0]), dec.model.quantized_blocks,
                                                                   ''.join(w.split('/')[0] for w in res[0][0][0][1])[:30]))

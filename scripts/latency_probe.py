#!/usr/bin/env python
"""Single-sentence decode (latency mode: float64 back end, few rows in flight) at the cfg-2 model size; prints the
per-call latency.  Used under ncu to look at the weight-streaming kernels."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import config, synth  # noqa: E402

root = tempfile.mkdtemp(prefix='jlm_lat_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
sents = synth.make_sentences(lexicon, 4, min_len=20, seed=100, vocab_size=50000)
config.set_root(root)
dec = jlm_b200.Decoder(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 0      # 0 auto (float64 for one sentence), 1 float64, 2 tensor cores
for s in sents[:2]:
    dec.decode_batch([s], topN=10, beam_width=10, backend=backend)
t0 = time.perf_counter()
for i in range(n):
    dec.decode_batch([sents[i % 4]], topN=10, beam_width=10, backend=backend)
print('single-sentence decode_batch (backend %d): %.3f ms per call (%d kana)'
      % (backend, (time.perf_counter() - t0) / n * 1e3, len(sents[0])))

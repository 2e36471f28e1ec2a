#!/usr/bin/env python
"""Tensor-core back end vs the float64 exact back end on a full-size lock-step batch (cfg 2 shape):
how many n-best lists / top-1 results are identical, and the worst score difference."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import config, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
root = tempfile.mkdtemp(prefix='jlm_par_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 321
sents = synth.make_sentences(lexicon, n, min_len=20, seed=seed, vocab_size=50000)
config.set_root(root)
dec = jlm_b200.Decoder(1)
a = dec.decode_batch(sents, topN=10, beam_width=10, backend=1)
gaps = sorted(min(abs(x[i + 1][0] - x[i][0]) for i in range(len(x) - 1)) for x in a if len(x) > 1)
print('sentences %d, smallest adjacent n-best gap %.3e (median %.3e)' % (n, gaps[0], gaps[len(gaps) // 2]))
for eps in ((0.0, 0.0, -1.0, 3e-5, 1e-4, 3e-4, 1e-3) if len(sys.argv) <= 2 else (0.0, 1e-4)):
    dec.model.set_guard(eps)
    import time
    t0 = time.perf_counter()
    b = dec.decode_batch(sents, topN=10, beam_width=10, backend=2)
    dt = time.perf_counter() - t0
    info = dec.last_info
    same = sum([w for _, w in x] == [w for _, w in y] for x, y in zip(a, b))
    top1 = sum(x[0][1] == y[0][1] for x, y in zip(a, b))
    worst = max(abs(p[0] - q[0]) for x, y in zip(a, b) for p, q in zip(x, y))
    print('guard eps %.1e: flagged %d / %d (pairs re-scored %d, sentences re-decoded %d), identical n-best %d, identical top-1 %d, '
          'worst |score diff| %.3e, min gap %.3e, wall %.1f ms'
          % (info.guard_eps, info.n_guard_flagged, n, info.n_guard_pairs, info.n_guard_rerun, same, top1, worst,
             info.guard_min_gap, dt * 1e3))

#!/usr/bin/env python
"""Where does the tensor-core back end's score error come from?  Decodes the same cfg-2 batch on the float64
back end and on the tensor-core back end with per-frame traces and reports, per lock-step frame, the error of
h, c, LSE, the transition logit (y = LSE - (score - parent score)) and the path score for beam entries that
are the same path on both back ends, plus the distribution of adjacent score gaps the prune decisions rest on.
Input for choosing the near-tie guard's bound."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import config, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
root = tempfile.mkdtemp(prefix='jlm_err_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
sents = synth.make_sentences(lexicon, n, min_len=20, seed=321, vocab_size=50000)
config.set_root(root)
dec = jlm_b200.Decoder(1)
dec.model.set_guard(0.0)      # raw tensor-core results
dec._want_trace = True
dec.decode_batch(sents, topN=10, beam_width=10, backend=1)
ex = dec._last_batch_trace
dec.decode_batch(sents, topN=10, beam_width=10, backend=2)
tc = dec._last_batch_trace

Tmax = max(len(f) for f in ex)
err = {k: [[] for _ in range(Tmax)] for k in ('h', 'c', 'lse', 'score', 'trans')}
signed = {k: [] for k in ('lse', 'trans', 'hsum')}
gaps = []
diff_err = []          # error of the DIFFERENCE between adjacent kept scores: what a rank decision depends on
diff_err_sib = []      # ... for pairs that extend the same parent path
mismatch_frames = 0
for s in range(n):
    for t, (a, b) in enumerate(zip(ex[s], tc[s])):
        same = (len(a['node']) == len(b['node']) and np.array_equal(a['node'], b['node'])
                and np.array_equal(a['parent_rank'], b['parent_rank']) and np.array_equal(a['parent_frame'], b['parent_frame']))
        sc = np.sort(a['score'])
        if len(sc) > 1:
            gaps.extend(np.diff(sc).tolist())
        if not same:
            mismatch_frames += 1
            continue
        err['score'][t].append(np.abs(a['score'] - b['score']).max())
        if len(a['score']) > 1:
            de = np.abs(np.diff(a['score']) - np.diff(b['score']))
            sib = (a['parent_frame'][1:] == a['parent_frame'][:-1]) & (a['parent_rank'][1:] == a['parent_rank'][:-1])
            diff_err.extend(de.tolist())
            diff_err_sib.extend(de[sib].tolist())
        if t < len(ex[s]) - 1:
            err['h'][t].append(np.abs(a['h'] - b['h']).max())
            err['c'][t].append(np.abs(a['c'] - b['c']).max())
            err['lse'][t].append(np.abs(a['lse'] - b['lse']).max())
            signed['lse'].extend((b['lse'] - a['lse']).tolist())
            signed['hsum'].extend((np.abs(b['h']).sum(axis=1) - np.abs(a['h']).sum(axis=1)).tolist())
        if t > 0:
            # transition cost of each kept path = score - parent's score
            for src, dst in ((a, 'ea'), (b, 'eb')):
                pass
            ta = a['score'] - np.array([ex[s][pf]['score'][pr] for pf, pr in zip(a['parent_frame'], a['parent_rank'])])
            tb = b['score'] - np.array([tc[s][pf]['score'][pr] for pf, pr in zip(b['parent_frame'], b['parent_rank'])])
            err['trans'][t].append(np.abs(ta - tb).max())
            signed['trans'].extend((tb - ta).tolist())
print('sentences %d, frames with different beams: %d' % (n, mismatch_frames))
print('frame   max|dh|    max|dc|   max|dLSE|  max|dtrans| max|dscore|  (max over sentences; mean in brackets)')
for t in range(Tmax):
    def f(k):
        v = err[k][t]
        return '%.2e(%.1e)' % (max(v), float(np.mean(v))) if v else '     -        '
    print('%3d  %s %s %s %s %s' % (t, f('h'), f('c'), f('lse'), f('trans'), f('score')))
g = np.sort(np.array(gaps))
print('adjacent kept-score gaps: n=%d' % len(g))
for thr in (1e-6, 1e-5, 3e-5, 1e-4, 3e-4, 1e-3, 3e-3):
    print('  gap < %.0e: %d (%.3f %% of gaps)' % (thr, int((g < thr).sum()), 100.0 * (g < thr).mean()))
print('  exact zero gaps: %d' % int((g == 0).sum()))
for k, v in signed.items():
    v = np.array(v)
    print('signed tc - exact, %-6s: mean %+.3e  std %.3e  min %+.3e  max %+.3e' % (k, v.mean(), v.std(), v.min(), v.max()))
for name, v in (('all adjacent pairs', diff_err), ('same-parent pairs', diff_err_sib)):
    v = np.sort(np.array(v))
    print('error of adjacent score DIFFERENCES, %s: n=%d  median %.2e  99%% %.2e  99.9%% %.2e  max %.2e'
          % (name, len(v), v[len(v) // 2], v[int(len(v) * 0.99)], v[int(len(v) * 0.999)], v[-1]))

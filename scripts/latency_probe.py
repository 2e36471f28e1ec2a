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

# the same through the batch entry points without timers (what jlm_decode_batch does): the cooperative single-sentence
# kernel (k_single_f64) takes sentences decoded alone on the float64 back end unless per-bucket timers were asked for
from jlm_b200 import lattice, _lib  # noqa: E402
nlex = dec._native()
for timers in (False, True):
    packs = [lattice.NativeLattices(nlex, [s], _lib.DECODE_FULL, None) for s in sents]
    outs = [dec._run(p, _lib.DECODE_FULL, 10, 10, backend or _lib.BACKEND_EXACT, timers) for p in packs]
    t0 = time.perf_counter()
    for i in range(n):
        dec._run(packs[i % 4], _lib.DECODE_FULL, 10, 10, backend or _lib.BACKEND_EXACT, timers)
    print('upload + run + fetch, timers=%s: %.3f ms per call; %d launches per call'
          % (timers, (time.perf_counter() - t0) / n * 1e3, dec.last_info.kernel_launches))
    if timers is False:
        first = outs
    else:
        same = all([w for _, w in a[0]] == [w for _, w in b[0]] for a, b in zip(first, outs))
        worst = max(abs(p[0] - q[0]) for a, b in zip(first, outs) for p, q in zip(a[0], b[0]))
        print('single-kernel path vs per-frame launches: identical n-best %s, worst |score diff| %.3e' % (same, worst))

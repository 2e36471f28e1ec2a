#!/usr/bin/env python
"""Where the end-to-end time of one 1024-sentence call goes (GPU box): lattice build, plan + H2D
(jlm_batch_upload), device run, D2H + unpack (jlm_batch_fetch)."""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import _lib, config, lattice, synth  # noqa: E402

root = tempfile.mkdtemp(prefix='jlm_e2e_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
sents = synth.make_sentences(lexicon, 1024, min_len=20, seed=100, vocab_size=50000)
config.set_root(root)
dec = jlm_b200.Decoder(1)
lib, hdl, nlex = dec._lib, dec.model._handle, dec._native()
S, TOPN, BEAM = len(sents), 10, 10
max_len = max(len(s) for s in sents) + 1
scores = np.empty((S, TOPN)); n_paths = np.empty(S, dtype=np.int32)
path_len = np.empty((S, TOPN), dtype=np.int32); path_nodes = np.zeros((S, TOPN, max_len), dtype=np.int32)
nb = _lib.NBest()
nb.top_n, nb.max_len = TOPN, max_len
nb.scores, nb.n_paths = _lib.ptr(scores, C.c_double), _lib.ptr(n_paths, C.c_int32)
nb.path_len, nb.path_nodes = _lib.ptr(path_len, C.c_int32), _lib.ptr(path_nodes, C.c_int32)
acc = np.zeros(4)
for it in range(8):
    t0 = time.perf_counter()
    packed = lattice.NativeLattices(nlex, sents)
    lb = packed.c_struct()
    t1 = time.perf_counter()
    batch = C.c_void_p()
    _lib.check(lib.jlm_batch_upload(hdl, C.byref(lb), BEAM, TOPN, 0, 2, C.byref(batch)))
    t2 = time.perf_counter()
    _lib.check(lib.jlm_batch_run(batch))
    _lib.check(lib.jlm_synchronize(hdl))
    t3 = time.perf_counter()
    _lib.check(lib.jlm_batch_fetch(batch, C.byref(nb)))
    t4 = time.perf_counter()
    _lib.check(lib.jlm_batch_destroy(batch))
    if it >= 3:
        acc += [t1 - t0, t2 - t1, t3 - t2, t4 - t3]
print('ms per call: lattice %.2f  upload(plan+H2D) %.2f  run %.2f  fetch %.2f' % tuple(acc / 5 * 1e3))

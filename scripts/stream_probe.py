#!/usr/bin/env python
"""Host-side timeline of the streaming e2e path (jlm_decode_texts_submit / _collect): per-call wall time of
submit and collect at depth 1 and 2, alternating, on the cfg-2 bench workload."""
import ctypes as C
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import _lib, config, synth  # noqa: E402

root = tempfile.mkdtemp(prefix='jlm_e2e_')
cfg, weights, lexicon, reading = synth.make_experiment(root, 1, 50000, 512, 256, 'tied', seed=0)
S = int(os.environ.get('SENTS', '1024'))
sents = synth.make_sentences(lexicon, S, min_len=20, seed=100, vocab_size=50000)
config.set_root(root)
dec = jlm_b200.Decoder(1)
lib, hdl, nlex = dec._lib, dec.model._handle, dec._native()
TOPN = BEAM = 10
if os.environ.get('PROBE_TORCH_STREAM'):
    import torch
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream()
    _lib.check(lib.jlm_set_stream(hdl, C.c_void_p(stream.cuda_stream)))
    print('using torch stream', stream.cuda_stream)
if os.environ.get('PROBE_SMI'):
    import subprocess
    smi = subprocess.Popen(['nvidia-smi', '-i', '0', '--query-gpu=clocks.sm', '--format=csv,noheader,nounits', '-lms', '100'],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lens = np.array([len(t) for t in sents], dtype=np.int64)
tptr = np.zeros(S + 1, dtype=np.int64)
np.cumsum(lens, out=tptr[1:])
cps = np.frombuffer(''.join(sents).encode('utf-32-le'), dtype=np.uint32)
tmax = int(lens.max()) + 1
sc = np.empty((S, TOPN)); npth = np.empty(S, dtype=np.int32); ln = np.empty((S, TOPN), dtype=np.int32)
ent = np.zeros((S, TOPN, tmax), dtype=np.int32); st = np.zeros((S, TOPN, tmax), dtype=np.int32)
tnb = _lib.TextNBest()
tnb.top_n, tnb.max_len = TOPN, tmax
tnb.scores, tnb.n_paths, tnb.path_len = _lib.ptr(sc, C.c_double), _lib.ptr(npth, C.c_int32), _lib.ptr(ln, C.c_int32)
tnb.path_entry, tnb.path_start = _lib.ptr(ent, C.c_int32), _lib.ptr(st, C.c_int32)
chars = int(lens.sum())


def submit():
    job = C.c_void_p()
    _lib.check(lib.jlm_decode_texts_submit(hdl, nlex.handle, S, _lib.ptr(tptr, C.c_int64), _lib.ptr(cps, C.c_uint32),
                                           BEAM, TOPN, 0, 0, None, 2, 0, 0, C.byref(job)))
    return job


def run(n, depth):
    ts, tc = [], []
    lib.jlm_synchronize(hdl)
    t0 = time.perf_counter()
    q = []
    for _ in range(n):
        a = time.perf_counter()
        q.append(submit())
        b = time.perf_counter()
        ts.append(b - a)
        if len(q) >= depth:
            _lib.check(lib.jlm_decode_texts_collect(q.pop(0), C.byref(tnb), None))
            tc.append(time.perf_counter() - b)
    while q:
        b = time.perf_counter()
        _lib.check(lib.jlm_decode_texts_collect(q.pop(0), C.byref(tnb), None))
        tc.append(time.perf_counter() - b)
    dt = time.perf_counter() - t0
    return dt / n * 1e3, np.array(ts) * 1e3, np.array(tc) * 1e3


for _ in range(3):
    run(2, 1)
for rep in range(2):
    for depth in (2, 1):
        ms, ts, tc = run(12, depth)
        print('depth %d: %.2f ms/batch (%.2f M chars/s)  submit ms %s  collect ms %s' % (
            depth, ms, chars / ms / 1e3, np.round(ts, 1).tolist(), np.round(tc, 1).tolist()), flush=True)

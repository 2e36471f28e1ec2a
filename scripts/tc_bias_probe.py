#!/usr/bin/env python
"""Signed error of the split-fp16 tcgen05 GEMM (jlm_tc_gemm_selftest, fp32 accumulation in TMEM) against float64:
is the accumulation error a bias proportional to the value (round-toward-zero accumulate) or zero-mean noise?
Regresses err = C_tc - C_ref on C_ref per K; reports beta (slope), the residual after removing beta*C and the raw rms."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import _lib, config, synth  # noqa: E402

root = tempfile.mkdtemp(prefix='jlm_bias_')
synth.make_experiment(root, 1, 1000, 64, 32, 'tied', seed=0)
config.set_root(root)
dec = jlm_b200.Decoder(1)
rng = np.random.default_rng(0)
for K in (64, 128, 256, 512, 768, 1280):
    for name, gen in (('gauss', lambda s: rng.normal(0, 1, size=s)), ('positive', lambda s: np.abs(rng.normal(0, 1, size=s)))):
        M, N = 512, 1024
        A = gen((M, K)).astype(np.float32)
        B = gen((N, K)).astype(np.float32)
        out = np.zeros((M, N), dtype=np.float32)
        ms = C.c_float(0)
        _lib.check(dec._lib.jlm_tc_gemm_selftest(dec.model._handle, _lib.ptr(A, C.c_float), _lib.ptr(B, C.c_float),
                                                 M, N, K, _lib.ptr(out, C.c_float), C.byref(ms)))
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        err = out.astype(np.float64) - ref
        beta = float((err * ref).sum() / (ref * ref).sum())
        resid = err - beta * ref
        rms = float(np.sqrt((ref * ref).mean()))
        # fp32 rounding of the stored result alone contributes ~ 2^-25 rms relative
        print('K=%4d %-8s rms|C| %8.2f  beta %+.3e  rms err/rms %.3e  rms resid/rms %.3e  max|err|/rms %.3e'
              % (K, name, rms, beta, np.sqrt((err ** 2).mean()) / rms, np.sqrt((resid ** 2).mean()) / rms,
                 np.abs(err).max() / rms))

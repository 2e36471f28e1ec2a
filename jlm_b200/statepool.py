"""StatePool: LSTM_Model.predict_with_context (reference decoder/model.py:195-198) with the (hidden, cell) pairs kept
on the device and named by slot - the C ABI's jlm_pool_* entry points.  A host-driven search sends word indices
and slot numbers down and gets one float64 (-log p) per asked (state, word) pair back."""
import ctypes as C

import numpy as np

from . import _lib


class StatePool(object):
    def __init__(self, model, capacity):
        self._lib = _lib.load()
        self._model = model                     # keeps the handle alive
        self._p = C.c_void_p()
        self.capacity = int(capacity)
        _lib.check(self._lib.jlm_pool_create(model._handle, self.capacity, C.byref(self._p)))
        self.used = 0

    def __del__(self):
        try:
            if self._p:
                self._lib.jlm_pool_destroy(self._p)
                self._p = None
        except Exception:
            pass

    def reset(self):
        _lib.check(self._lib.jlm_pool_reset(self._p))
        self.used = 0

    def step(self, src, index):
        """One batched LM step: row k continues the state in slot src[k] (-1: zero state) with word index[k].
        Returns the slots of the new states (consecutive)."""
        src = np.ascontiguousarray(src, dtype=np.int32).reshape(-1)
        index = np.ascontiguousarray(index, dtype=np.int32).reshape(-1)
        if src.shape != index.shape:
            raise ValueError('src and index differ in length')
        first = C.c_int64(0)
        _lib.check(self._lib.jlm_pool_step(self._p, len(src), _lib.ptr(src, C.c_int32), _lib.ptr(index, C.c_int32),
                                           C.byref(first)))
        self.used = first.value + len(src)
        return np.arange(first.value, first.value + len(src), dtype=np.int64)

    def nll(self, slots, cols):
        """-log p(cols[k] | state slots[k]) as float64 (decoder.py:43-49; -y for self-normalised models)."""
        slots = np.ascontiguousarray(slots, dtype=np.int32).reshape(-1)
        cols = np.ascontiguousarray(cols, dtype=np.int32).reshape(-1)
        if slots.shape != cols.shape:
            raise ValueError('slots and cols differ in length')
        out = np.empty(len(slots))
        if len(slots):
            _lib.check(self._lib.jlm_pool_nll(self._p, len(slots), _lib.ptr(slots, C.c_int32), _lib.ptr(cols, C.c_int32),
                                              _lib.ptr(out, C.c_double)))
        return out

    def state(self, slot, count=1):
        H = self._model.hidden_size
        h = np.empty((count, H))
        c = np.empty((count, H))
        _lib.check(self._lib.jlm_pool_get_state(self._p, int(slot), int(count), _lib.ptr(h, C.c_double),
                                                _lib.ptr(c, C.c_double)))
        return h, c

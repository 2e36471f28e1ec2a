"""LSTM_Model: drop-in for the reference's numpy language model (decoder/model.py:35-206), with the
arithmetic on a B200 through libjlm_b200.so.

Same constructor, attributes (config, weights, hidden, cell, hidden_size, embed_size,
share_embedding, blocks, v_tables) and methods (predict, predict_with_context, project), same
return shapes and dtypes (float64), same quirks where they are observable.
There is no CPU code path: constructing the model without the CUDA library or a GPU raises.
"""
import ctypes as C
import os
import pickle
import sys

import numpy as np

from . import _lib, config


def _f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class LSTM_Model(object):
    """The B200 implementation of the NN language model (reference: decoder/model.py:35)."""

    def __init__(self, experiment_id=0, comp=0, device=0, config_dict=None, weights=None):
        # decoder/model.py:37-45
        self.config = config_dict if config_dict is not None else config.load_config(experiment_id)
        self.weights = weights if weights is not None else self._load_model(experiment_id, comp)
        self.embed_size = self.config['embed_size']
        self.hidden_size = self.config['hidden_size']
        self.share_embedding = self.config['share_embedding']
        self.hidden = np.zeros((1, self.hidden_size))
        self.cell = np.zeros((1, self.hidden_size))
        self.device = device
        cfg, w = self.config, self.weights
        V = int(w['b2'].shape[0])
        segs = [(int(sz), int(s), V if e is None else int(e)) for sz, s, e in cfg.get('embedding_seg') or []]

        if cfg.get('D_softmax'):
            # decoder/model.py:47-54: block-diagonal float64 input table [V, sum e_i]
            self.blocks = w['LM']
            self.embed_size = sum(x[0] for x in cfg['embedding_seg'])
            LM = np.zeros((V, self.embed_size))
            col = 0
            for i, (size, s, e) in enumerate(segs):
                LM[s:e, col:col + size] = self.blocks[i]
                col += size
            w['LM'] = LM
        if cfg.get('V_table'):
            # decoder/model.py:56-71: LM = concat(LM0, LM1.VT1, ...) (float32 products, as the reference)
            self.blocks, self.v_tables, emb = [], [], []
            for i, _ in enumerate(segs):
                blk = w['LM{}'.format(i)]
                self.blocks.append(blk)
                if i:
                    vt = w['VT{}'.format(i)]
                    self.v_tables.append(vt)
                    emb.append(np.dot(blk, vt))
                else:
                    self.v_tables.append(None)
                    emb.append(blk)
            w['LM'] = np.concatenate(emb, axis=0)

        # ---- hand the float32 weights to the device library (jlm_create) ----
        c = _lib.Config()
        c.vocab_size, c.hidden_size = V, int(self.hidden_size)
        c.self_norm = 1 if cfg.get('self_norm') else 0
        keep = []

        def fp(a):
            a = _f32c(a)
            keep.append(a)
            return _lib.ptr(a, C.c_float)

        cw = _lib.Weights()
        for g, name in enumerate('ifog'):
            cw.HM[g] = fp(w['HM' + name])
            cw.IM[g] = fp(w['IM' + name])
            cw.b[g] = fp(w['b' + name])
        cw.b2 = fp(w['b2'])
        cw.LM_in = fp(w['LM'])
        c.input_embed = int(np.asarray(w['LM']).shape[1])
        if not self.share_embedding:
            c.proj_mode = _lib.PROJ_UNTIED
            c.n_seg = 1
            c.seg_width[0], c.seg_start[0], c.seg_end[0] = int(self.hidden_size), 0, V
            cw.UM = fp(w['UM'])
        elif cfg.get('D_softmax') or cfg.get('V_table'):
            c.proj_mode = _lib.PROJ_DSOFTMAX if cfg.get('D_softmax') else _lib.PROJ_DSOFTMAX_STAR
            if len(segs) > _lib.MAX_SEGMENTS:
                raise ValueError('at most %d embedding segments are supported' % _lib.MAX_SEGMENTS)
            c.n_seg = len(segs)
            for i, (sz, s, e) in enumerate(segs):
                c.seg_width[i], c.seg_start[i], c.seg_end[i] = sz, s, e
                cw.seg_LM[i] = fp(self.blocks[i])
                if cfg.get('V_table') and i:
                    cw.seg_VT[i] = fp(self.v_tables[i])
            cw.PM = fp(w['PM'])
        else:
            c.proj_mode = _lib.PROJ_TIED
            c.n_seg = 1
            c.seg_width[0], c.seg_start[0], c.seg_end[0] = int(np.asarray(w['LM']).shape[1]), 0, V
            cw.seg_LM[0] = cw.LM_in
            cw.PM = fp(w['PM'])
        self._segs = segs
        self._V = V
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        _lib.check(self._lib.jlm_create(C.byref(c), C.byref(cw), int(device), C.byref(self._handle)))
        del keep
        self.quantized_blocks = self._attach_codes(experiment_id, comp) if (comp and weights is None) else []

    def _attach_codes(self, experiment_id, comp):
        """Hands the k-means (code, codebook) form of the output blocks to the device when the reference's
        comp dump is present (train/comp.py:79-80: weights/comp_N/lstm_weights_comp_dump.pkl).  The codes
        must decode to the floats already loaded (jlm_set_quantized_block verifies), so results are
        unchanged; only the bytes streamed per LM step by the few-rows-in-flight path shrink 4x."""
        path = os.path.join(config.experiment_path, str(experiment_id), 'weights', 'comp_{}'.format(comp),
                            'lstm_weights_comp_dump.pkl')
        if comp > 8 or not os.path.exists(path):
            return []
        with open(path, 'rb') as f:
            dump = pickle.load(f)
        if not self.share_embedding:
            keys = [('UM', True)]
        elif self.config.get('V_table'):
            keys = [('LM{}'.format(i), False) for i in range(len(self._segs))]
        elif self.config.get('D_softmax'):
            return []                      # comp.py cannot compress the block list of a D-softmax dump
        else:
            keys = [('LM', False)]
        done = []
        for seg, (key, transposed) in enumerate(keys):
            if key not in dump:
                continue
            code, codebook = dump[key]
            code = np.asarray(code)
            if code.dtype != np.uint8:
                continue
            code = np.ascontiguousarray(code.T if transposed else code)
            cb = np.ascontiguousarray(np.asarray(codebook, dtype=np.float32).reshape(-1))
            _lib.check(self._lib.jlm_set_quantized_block(self._handle, seg, _lib.ptr(code, C.c_uint8),
                                                         _lib.ptr(cb, C.c_float), int(cb.shape[0])))
            done.append(key)
        return done

    def __del__(self):
        h = getattr(self, '_handle', None)
        if h is not None and h.value:
            self._lib.jlm_destroy(h)
            self._handle = None

    def set_guard(self, eps, verify=True, scope='output'):
        """Near-tie guard of the tensor-core back end (jlm_set_guard): rank decisions that rest on a score gap below
        `eps` are re-scored in float64 and, if contradicted, their sentence is re-decoded in float64.  eps = 0
        switches it off, a negative value restores the default; verify=False skips the re-scoring tier and
        re-decodes every flagged sentence; scope='output' guards the decisions that can change the returned n-best
        (kept/rejected boundary of every frame, order of the last frame), scope='all' every adjacent pair of every
        frame (the per-frame traces too)."""
        if scope not in ('output', 'all'):
            raise ValueError("scope must be 'output' or 'all'")
        _lib.check(self._lib.jlm_set_guard(self._handle, float(eps)))
        _lib.check(self._lib.jlm_set_guard_verify(self._handle, 1 if verify else 0))
        _lib.check(self._lib.jlm_set_guard_scope(self._handle, 1 if scope == 'all' else 0))

    def _load_model(self, experiment_id=0, comp=0):
        # decoder/model.py:73-104 (comp>0 loads the already-decoded k-means pickle, quirk 11)
        name = 'lstm_weights_comp_{}.pkl'.format(comp) if comp else 'lstm_weights.pkl'
        wdir = os.path.join(config.experiment_path, str(experiment_id), 'weights')
        path = os.path.join(wdir, name)
        if not comp and not os.path.exists(path) and os.path.exists(os.path.join(wdir, 'b2.txt')):
            # only the per-parameter text dumps of `weights.py --verbose` survive (train/weights.py:77-87)
            from . import weights_io
            return weights_io.load_text_weights(wdir, self.config)
        with open(path, 'rb') as f:
            return pickle.load(f)

    # ------------------------------------------------------------------------------------------
    def _vocab_columns(self, vocab):
        """(cols, bias_idx) reproducing the reference's column order for a vocab subset.
        Segmented models emit columns segment-major while b2[vocab] stays in list order
        (decoder/model.py:152-158,168-179)."""
        vocab = [int(v) for v in vocab]
        if not self.share_embedding:
            # decoder/model.py:189: UM[vocab] indexes rows of [H,V] (quirk 2)
            if max(vocab) >= self.hidden_size or min(vocab) < -self.hidden_size:
                raise IndexError('index {} is out of bounds for axis 0 with size {}'.format(max(vocab), self.hidden_size))
            raise ValueError('shapes ({},{}) and ({},{}) not aligned'.format(0, self.hidden_size, len(vocab), self._V))
        if self.config.get('D_softmax') or self.config.get('V_table'):
            cols = []
            for _, s, e in [(sz, s, (sys.maxsize if raw[2] is None else e))
                            for (sz, s, e), raw in zip(self._segs, self.config['embedding_seg'])]:
                cols += [v for v in vocab if s <= v < e]
            if len(cols) != len(vocab):
                raise ValueError('operands could not be broadcast together')
        else:
            cols = vocab
        for v in cols:
            if not 0 <= v < self._V:
                raise IndexError('index {} is out of bounds for axis 0 with size {}'.format(v, self._V))
        return np.array(cols, dtype=np.int32), np.array(vocab, dtype=np.int32)

    def _state(self, a, B):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[None, :]
        if a.shape[0] == 1 and B > 1:
            a = np.repeat(a, B, axis=0)      # the reference broadcasts a [1,H] state over the batch
        if a.shape != (B, self.hidden_size):
            raise ValueError('state shape {} does not match batch {}'.format(a.shape, B))
        return np.ascontiguousarray(a)

    def predict(self, index, vocab=None, reset=False):
        """decoder/model.py:106-123 -> (pred [B,N], y [B,N], seconds_lstm, seconds_softmax)."""
        if reset:
            self.hidden = np.zeros(shape=self.hidden.shape)
            self.cell = np.zeros(shape=self.cell.shape)
        idx = np.ascontiguousarray(np.asarray(index, dtype=np.int32).reshape(-1))
        B = idx.shape[0]
        h_in, c_in = self._state(self.hidden, B), self._state(self.cell, B)
        cols = bias = None
        N = self._V
        if vocab:
            cols, bias = self._vocab_columns(vocab)
            N = len(cols)
        pred = np.empty((B, N))
        y = np.empty((B, N))
        h_out = np.empty((B, self.hidden_size))
        c_out = np.empty((B, self.hidden_size))
        t1, t2 = C.c_float(0), C.c_float(0)
        d = C.c_double
        _lib.check(self._lib.jlm_predict(self._handle, _lib.ptr(idx, C.c_int32), _lib.ptr(h_in, d), _lib.ptr(c_in, d),
                                         B, _lib.ptr(cols, C.c_int32), _lib.ptr(bias, C.c_int32), N,
                                         _lib.ptr(pred, d), _lib.ptr(y, d), _lib.ptr(h_out, d), _lib.ptr(c_out, d),
                                         C.byref(t1), C.byref(t2)))
        self.hidden, self.cell = h_out, c_out
        return pred, y, t1.value * 1e-3, t2.value * 1e-3

    def _lstm_cell(self, index):
        """decoder/model.py:125-139 (updates self.hidden / self.cell in place of the instance)."""
        idx = np.ascontiguousarray(np.asarray(index, dtype=np.int32).reshape(-1))
        B = idx.shape[0]
        h_in, c_in = self._state(self.hidden, B), self._state(self.cell, B)
        h_out = np.empty((B, self.hidden_size))
        c_out = np.empty((B, self.hidden_size))
        d = C.c_double
        _lib.check(self._lib.jlm_lstm_step(self._handle, _lib.ptr(idx, C.c_int32), _lib.ptr(h_in, d),
                                           _lib.ptr(c_in, d), B, _lib.ptr(h_out, d), _lib.ptr(c_out, d)))
        self.hidden, self.cell = h_out, c_out

    def project(self, hidden, vocab=None):
        """decoder/model.py:141-193 -> y [B, V or len(vocab)] float64."""
        hidden = np.asarray(hidden, dtype=np.float64)
        if hidden.ndim == 1:
            hidden = hidden[None, :]
        hidden = np.ascontiguousarray(hidden)
        B = hidden.shape[0]
        cols = bias = None
        N = self._V
        if vocab:
            cols, bias = self._vocab_columns(vocab)
            N = len(cols)
        y = np.empty((B, N))
        d = C.c_double
        _lib.check(self._lib.jlm_project(self._handle, _lib.ptr(hidden, d), B, _lib.ptr(cols, C.c_int32),
                                         _lib.ptr(bias, C.c_int32), N, _lib.ptr(y, d)))
        return y

    def predict_with_context(self, index, hidden, cell, vocab=None):
        """decoder/model.py:195-198"""
        self.hidden = hidden
        self.cell = cell
        return self.predict(index, vocab), self.hidden, self.cell

    def evaluate(self, start, inputs):
        """decoder/model.py:200-206, with the tuple indexing the reference gets wrong (quirk 1) fixed:
        returns the per-step negative log probabilities of `inputs` after `start`."""
        probs = []
        pred = self.predict([start], vocab=None, reset=True)[0]
        for inp in inputs:
            probs.append(pred[0, inp])
            pred = self.predict([inp])[0]
        return [-np.log(p) for p in probs]

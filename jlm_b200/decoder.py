"""Decoder: drop-in for the reference's Viterbi / beam-search kana->kanji decoder
(decoder/decoder.py:54-241), running on a B200 through libjlm_b200.so.

decode() keeps the reference signature and return value; decode_batch() decodes many independent
sentences in lock-step (the throughput path).  The lattice is built on the host exactly as the
reference does (order matters: ties are broken by enumeration order), everything after that - LM
steps, softmax statistics, expand/score/prune, back-trace - runs on the device.
"""
import ctypes as C
import os
import pickle

import numpy as np

from . import _lib, config, lattice
from .model import LSTM_Model
from .vocab import Vocab


class Decoder(object):
    dynamic = False

    def __init__(self, experiment_id=0, comp=0, device=0):
        # decoder/decoder.py:55-69
        self.config = config.load_config(experiment_id)
        self._load_vocab()
        with open(os.path.join(config.root_path, 'data', 'lexicon.pkl'), 'rb') as f:
            self.full_lexicon = pickle.load(f)
        with open(os.path.join(config.root_path, 'data', 'reading_dict.pkl'), 'rb') as f:
            self.full_reading_dict = pickle.load(f)
        if bool(self.config.get('char_rnn')) != bool(getattr(self, 'char_rnn', False)):
            # the reference picks the class from config['char_rnn'] (decoder.py:345-348, eval.py:43-44)
            raise ValueError('config char_rnn={} needs {}'.format(self.config.get('char_rnn'),
                                                                  'CharRNNDecoder' if self.config.get('char_rnn') else 'Decoder'))
        self.model = LSTM_Model(experiment_id, comp, device=device)
        self.lattice_vocab = None
        self.backward_lookup = None
        self.perf_sen = 0
        self.perf_log_lstm = []
        self.perf_log_softmax = []
        # perf_log_lstm / perf_log_softmax (decoder.py:211-212) need CUDA events between the kernels of every frame, and
        # with those a sentence decoded alone goes through ~8 launches per frame instead of the one cooperative kernel
        # (k_single_f64): set perf_timers = False where the two perf logs are not read (interactive, one sentence at a time)
        self.perf_timers = True
        self._builder = lattice.LatticeBuilder(self.w2i, self.full_lexicon, self.full_reading_dict)
        self._native_lexicon = None
        self._lib = _lib.load()
        self.last_info = None

    def _load_vocab(self):
        # decoder/decoder.py:71-74
        self.vocab = Vocab(self.config['vocab_size'])
        self.i2w = self.vocab.i2w
        self.w2i = self.vocab.w2i

    def _check_oov(self, word):
        # decoder/decoder.py:76-77
        return word not in self.w2i

    def _build_lattice(self, input, vocab_select=False, samples=0, top_sampling=False, random_sampling=False):
        """decoder/decoder.py:79-135; returns frames (list of node tuples) and sets lattice_vocab."""
        frames = self._builder.build(input)
        if vocab_select:
            self._build_lattice_vocab(frames, samples, top_sampling, random_sampling)
        return frames

    def _build_lattice_vocab(self, frames, samples=0, top_sampling=False, random_sampling=False):
        # decoder/decoder.py:137-151
        self.lattice_vocab = lattice.static_vocab(frames, len(self.w2i), samples, top_sampling, random_sampling)

    # ------------------------------------------------------------------------------------------
    def _pack(self, all_frames, vocab_state):
        """vocab_state: per sentence None / list (static) -> PackedLattices + decode mode."""
        if vocab_state[0] is None:
            return lattice.PackedLattices(all_frames), _lib.DECODE_FULL
        for frames, lv in zip(all_frames, vocab_state):
            have = set(lv)
            for fr in frames:
                for n in fr:
                    if n[1] not in have:
                        # the reference fails in list.index (decoder.py:179-180, quirk 6)
                        raise ValueError('{} is not in list'.format(n[1]))
        return lattice.PackedLattices(all_frames, vocab_lists=vocab_state), _lib.DECODE_STATIC_VOCAB

    @staticmethod
    def _beam_args(topN, beam_width):
        """(beam_width, top) for the C ABI.  beam_width=None is the reference's unpruned search (decoder.py:227-229
        skipped): every candidate is kept in enumeration order, nothing is sorted; the library refuses batches
        whose path count explodes (JLM_MAX_UNPRUNED_PATHS)."""
        if beam_width is None:
            return _lib.BEAM_UNLIMITED, max(1, int(topN))
        beam_width = int(beam_width)
        if beam_width < 1:
            # frame[i][:0] is empty in the reference, the next frame has nothing to extend and frame 0 is lost
            raise ValueError('beam_width must be >= 1 or None')
        return beam_width, max(1, min(int(topN), beam_width))

    def _run(self, packed, mode, topN, beam_width, backend, timers):
        lib, h = self._lib, self.model._handle
        beam_width, top = self._beam_args(topN, beam_width)
        lb = packed.c_struct()
        batch = C.c_void_p()
        _lib.check(lib.jlm_batch_upload(h, C.byref(lb), int(beam_width), top, mode, backend, C.byref(batch)))
        try:
            if timers:
                _lib.check(lib.jlm_batch_enable_timers(batch, 1))
            _lib.check(lib.jlm_batch_run(batch))
            S = packed.n_sent
            max_len = int(packed.sent_len.max()) + 1
            scores = np.empty((S, top))
            n_paths = np.empty(S, dtype=np.int32)
            path_len = np.empty((S, top), dtype=np.int32)
            path_nodes = np.zeros((S, top, max_len), dtype=np.int32)
            nb = _lib.NBest()
            nb.top_n, nb.max_len = top, max_len
            nb.scores = _lib.ptr(scores, C.c_double)
            nb.n_paths = _lib.ptr(n_paths, C.c_int32)
            nb.path_len = _lib.ptr(path_len, C.c_int32)
            nb.path_nodes = _lib.ptr(path_nodes, C.c_int32)
            _lib.check(lib.jlm_batch_fetch(batch, C.byref(nb)))
            info = _lib.BatchInfo()
            _lib.check(lib.jlm_batch_get_info(batch, C.byref(info)))
            self.last_info = info
            self._last_batch_trace = None
            if getattr(self, '_want_trace', False):
                self._last_batch_trace = self._collect_trace(batch, packed, int(info.beam_width))
        finally:
            lib.jlm_batch_destroy(batch)
        out = []
        for s in range(S):
            res = []
            for k in range(int(n_paths[s])):
                ws = packed.path_words(s, path_nodes[s, k, :path_len[s, k]].tolist())
                res.append((float(scores[s, k]), [w for w in ws if w != '<eos>']))   # decoder.py:237
            out.append(res[:topN])
        return out

    def _submit_texts(self, texts, mode, extra, topN, beam_width, backend, n_chunks=0):
        """jlm_decode_texts_submit: lattice build, plan, H2D and every frame enqueued; returns a pending job
        (dict) without waiting for the device."""
        lib, h = self._lib, self.model._handle
        beam_width, top = self._beam_args(topN, beam_width)
        nlex = self._native()
        S = len(texts)
        lens = np.fromiter((len(t) for t in texts), dtype=np.int64, count=S)
        tptr = np.zeros(S + 1, dtype=np.int64)
        np.cumsum(lens, out=tptr[1:])
        joined = ''.join(texts)
        cps = np.frombuffer(joined.encode('utf-32-le'), dtype=np.uint32) if joined else np.zeros(1, dtype=np.uint32)
        n_extra = 0
        if extra is not None:
            extra = np.ascontiguousarray(extra, dtype=np.int32).reshape(S, -1)
            n_extra = extra.shape[1]
        job = C.c_void_p()
        _lib.check(lib.jlm_decode_texts_submit(h, nlex.handle, S, _lib.ptr(tptr, C.c_int64), _lib.ptr(cps, C.c_uint32),
                                               int(beam_width), top, int(mode), n_extra,
                                               _lib.ptr(extra, C.c_int32) if n_extra else None, int(backend),
                                               int(n_chunks), 1 if self.perf_timers else 0, C.byref(job)))
        return {'job': job, 'texts': texts, 'top': top, 'topN': topN, 'max_len': int(lens.max()) + 1}

    def _collect_texts_arrays(self, pending):
        """jlm_decode_texts_collect: waits for that job only; the n-best block as arrays
        {'scores' [S,top] f64, 'n_paths' [S], 'path_len' [S,top], 'path_entry' / 'path_start' [S,top,max_len] i32}
        (paths as lexicon entries: -1 '<eos>', -2 '<unk>' whose word is the kana at path_start)."""
        lib = self._lib
        texts, top, max_len = pending['texts'], pending['top'], pending['max_len']
        S = len(texts)
        scores = np.empty((S, top))
        n_paths = np.empty(S, dtype=np.int32)
        path_len = np.empty((S, top), dtype=np.int32)
        path_entry = np.zeros((S, top, max_len), dtype=np.int32)
        path_start = np.zeros((S, top, max_len), dtype=np.int32)
        nb = _lib.TextNBest()
        nb.top_n, nb.max_len = top, max_len
        nb.scores, nb.n_paths = _lib.ptr(scores, C.c_double), _lib.ptr(n_paths, C.c_int32)
        nb.path_len = _lib.ptr(path_len, C.c_int32)
        nb.path_entry, nb.path_start = _lib.ptr(path_entry, C.c_int32), _lib.ptr(path_start, C.c_int32)
        info = _lib.BatchInfo()
        job, pending['job'] = pending['job'], None       # collect consumes the job, also on error
        _lib.check(lib.jlm_decode_texts_collect(job, C.byref(nb), C.byref(info)))
        self.last_info = info
        return {'scores': scores, 'n_paths': n_paths, 'path_len': path_len, 'path_entry': path_entry,
                'path_start': path_start}

    def words_from_arrays(self, texts, arrays, topN=None):
        """n-best word lists [(neg_log_prob, [word, ...])] from the array form (decoder.py:237-241)."""
        nlex = self._native()
        if getattr(nlex, '_entry_words_obj', None) is None:
            nlex._entry_words_obj = np.array(list(nlex.entry_words) + [None, None], dtype=object)   # [-2], [-1] -> None
        ew = nlex._entry_words_obj
        scores, n_paths, path_len = arrays['scores'], arrays['n_paths'], arrays['path_len']
        path_entry, path_start = arrays['path_entry'], arrays['path_start']
        words = ew[path_entry]                      # one vectorised gather for the whole block (negative ids -> None)
        has_unk = (path_entry == -2).any(axis=2)
        sc_list, np_list, len_list = scores.tolist(), n_paths.tolist(), path_len.tolist()
        out = []
        for s, text in enumerate(texts):
            res = []
            for k in range(np_list[s]):
                n = len_list[s][k]
                ws = words[s, k, :n].tolist()
                if has_unk[s, k]:
                    # '<unk>' (-2) carries the raw kana of its frame (decoder.py:130)
                    ents, starts = path_entry[s, k, :n].tolist(), path_start[s, k, :n].tolist()
                    ws = [text[st] if e == -2 else w for w, e, st in zip(ws, ents, starts)]
                # '<eos>' (-1) is dropped (decoder.py:237): it is the first node of every path and maps to None
                res.append((sc_list[s][k], [w for w in ws if w is not None]))
            out.append(res[:topN] if topN is not None else res)
        return out

    def _collect_texts(self, pending):
        """jlm_decode_texts_collect: waits for that job only and turns its n-best block into word lists."""
        return self.words_from_arrays(pending['texts'], self._collect_texts_arrays(pending), pending['topN'])

    def decode_batch_arrays(self, inputs, topN=10, beam_width=10, backend=_lib.BACKEND_AUTO, **sampling):
        """decode_batch() returning the n-best block as arrays (see _collect_texts_arrays) instead of Python
        word lists: what a sharded caller exchanges between ranks (jlm_b200/shard.py).  Static full-softmax
        decoding for Decoder; DynamicDecoder overrides the mode."""
        inputs = list(inputs)
        mode, extra = self._array_mode(len(inputs), **sampling)
        arrays = self._collect_texts_arrays(self._submit_texts(inputs, mode, extra, topN, beam_width, backend))
        self._log_batch_perf(len(inputs), last_frame_stepped=not self.dynamic)
        return arrays

    def _array_mode(self, n_sent, vocab_select=False, samples=0, top_sampling=False, random_sampling=False):
        if vocab_select:
            return _lib.DECODE_STATIC_VOCAB, self._sample_ids(n_sent, samples, top_sampling, random_sampling)
        return _lib.DECODE_FULL, None

    def _run_texts(self, texts, mode, extra, topN, beam_width, backend, n_chunks=0):
        """One jlm_decode_texts call: kana strings in, n-best word lists out."""
        return self._collect_texts(self._submit_texts(texts, mode, extra, topN, beam_width, backend, n_chunks))

    def decode_stream(self, batches, topN=10, beam_width=10, backend=_lib.BACKEND_AUTO, depth=2):
        """decode_batch() over an iterable of sentence batches, as a generator of their results in order.
        Batch k+1 is submitted (lattices, plan, H2D, kernels enqueued) before batch k is collected, so the
        host work of one batch is hidden behind the device work of the previous one; `depth` batches are in
        flight at most.  Full-softmax decoding only (the mode decode_batch sends through jlm_decode_texts)."""
        if self.dynamic:
            raise ValueError('decode_stream: DynamicDecoder batches go through decode_batch')
        pending = []
        try:
            for texts in batches:
                texts = list(texts)
                if not texts:
                    raise ValueError('decode_stream: empty batch')
                pending.append(self._submit_texts(texts, _lib.DECODE_FULL, None, topN, beam_width, backend))
                while len(pending) >= max(1, int(depth)):
                    p = pending.pop(0)
                    out = self._collect_texts(p)
                    self._log_batch_perf(len(p['texts']))
                    yield out
            while pending:
                p = pending.pop(0)
                out = self._collect_texts(p)
                self._log_batch_perf(len(p['texts']))
                yield out
        finally:
            for p in pending:       # generator closed early or an error: wait for and drop what is in flight
                if p['job'] is not None:
                    self._lib.jlm_decode_texts_cancel(p['job'])

    def _collect_trace(self, batch, packed, W):
        """Per-frame pruned beams of every sentence (test / debugging aid)."""
        lib = self._lib
        H = self.model.hidden_size
        traces = []
        for s in range(packed.n_sent):
            T = int(packed.sent_len[s])
            n = (T + 1) * W
            count = np.zeros(T + 1, dtype=np.int32)
            score = np.zeros(n)
            pf = np.zeros(n, dtype=np.int32)
            pr = np.zeros(n, dtype=np.int32)
            node = np.zeros(n, dtype=np.int32)
            lse = np.zeros(n)
            hh = np.zeros((n, H))
            cc = np.zeros((n, H))
            d = C.c_double
            _lib.check(lib.jlm_batch_get_beams(batch, s, _lib.ptr(count, C.c_int32), _lib.ptr(score, d),
                                               _lib.ptr(pf, C.c_int32), _lib.ptr(pr, C.c_int32),
                                               _lib.ptr(node, C.c_int32), _lib.ptr(lse, d), _lib.ptr(hh, d),
                                               _lib.ptr(cc, d)))
            off = int(packed.node_off[s])
            frames = []
            for t in range(T + 1):
                sl = slice(t * W, t * W + int(count[t]))
                frames.append({'score': score[sl].copy(), 'parent_frame': pf[sl].copy(), 'parent_rank': pr[sl].copy(),
                               'node': node[sl] - off, 'lse': lse[sl].copy(), 'h': hh[sl].copy(), 'c': cc[sl].copy()})
            traces.append(frames)
        return traces

    # ------------------------------------------------------------------------------------------
    def decode(self, input, topN=10, beam_width=10, vocab_select=False, samples=0, top_sampling=False,
               random_sampling=False, backend=_lib.BACKEND_AUTO):
        """decoder/decoder.py:220-241 -> [(neg_log_prob, [word, ...])][:topN]."""
        frames = self._build_lattice(input, vocab_select=vocab_select, samples=samples, top_sampling=top_sampling,
                                     random_sampling=random_sampling)
        self.backward_lookup = lattice.to_backward_lookup(frames)
        # a vocabulary selected by an earlier call keeps being used (decoder.py:66,179; quirk 6)
        lv = self.lattice_vocab if self.lattice_vocab else None
        packed, mode = self._pack([frames], [lv])
        out = self._run(packed, mode, topN, beam_width, backend, timers=self.perf_timers)[0]
        info = self.last_info
        steps = max(int(info.n_steps), 1)
        if self.perf_timers:
            self.perf_log_lstm += [info.ms_lstm * 1e-3 / steps] * steps          # decoder.py:211-212
            self.perf_log_softmax += [info.ms_softmax * 1e-3 / steps] * steps
        self.perf_sen += 1
        return out

    def _native(self):
        if self._native_lexicon is None:
            self._native_lexicon = lattice.NativeLexicon(self.w2i, self.full_lexicon, self.full_reading_dict)
        return self._native_lexicon

    def _sample_ids(self, n_sent, samples, top_sampling, random_sampling):
        """The `samples` extra word ids of every sentence, drawn exactly as the reference does per
        sentence (decoder.py:144-149): np.random.randint from the global stream, or range(samples)."""
        if not samples or not (top_sampling or random_sampling):
            return None
        if random_sampling:
            return np.stack([np.random.randint(len(self.w2i), size=samples) for _ in range(n_sent)]).astype(np.int32)
        return np.tile(np.arange(samples, dtype=np.int32), (n_sent, 1))

    def decode_batch(self, inputs, topN=10, beam_width=10, vocab_select=False, samples=0, top_sampling=False,
                     random_sampling=False, backend=_lib.BACKEND_AUTO, native_lattice=True):
        """decode() for many independent sentences decoded in lock-step; returns one n-best list per input.
        The lattices are built by the native builder (jlm_lattice_build) unless native_lattice=False or a
        vocabulary left by an earlier decode() call has to be honoured (quirk 6)."""
        inputs = list(inputs)
        if not inputs:
            return []
        if native_lattice and (vocab_select or not self.lattice_vocab):
            mode = _lib.DECODE_STATIC_VOCAB if vocab_select else _lib.DECODE_FULL
            extra = self._sample_ids(len(inputs), samples, top_sampling, random_sampling) if vocab_select else None
            if not getattr(self, '_want_trace', False) and not vocab_select:
                out = self._run_texts(inputs, mode, extra, topN, beam_width, backend)
                self._log_batch_perf(len(inputs))
                return out
            packed = lattice.NativeLattices(self._native(), inputs, mode, extra)
            out = self._run(packed, mode, topN, beam_width, backend, timers=self.perf_timers)
            if vocab_select:
                self.lattice_vocab = packed.vocab_list(len(inputs) - 1)
            self._log_batch_perf(len(inputs))
            return out
        all_frames, vocabs = [], []
        for text in inputs:
            frames = self._build_lattice(text, vocab_select=vocab_select, samples=samples,
                                         top_sampling=top_sampling, random_sampling=random_sampling)
            all_frames.append(frames)
            vocabs.append(list(self.lattice_vocab) if self.lattice_vocab else None)
        packed, mode = self._pack(all_frames, vocabs)
        out = self._run(packed, mode, topN, beam_width, backend, timers=self.perf_timers)
        self._log_batch_perf(len(inputs))
        return out

    def _log_batch_perf(self, n_sent, last_frame_stepped=True):
        """perf_log_* for a lock-step batch: one entry per lock-step LM step (decoder.py:211-212), the
        CUDA-event time of that step for the whole batch."""
        info = self.last_info
        steps = max(int(info.n_steps) - (0 if last_frame_stepped else 1), 1)
        if self.perf_timers:
            self.perf_log_lstm += [info.ms_lstm * 1e-3 / steps] * steps
            self.perf_log_softmax += [info.ms_softmax * 1e-3 / steps] * steps
        self.perf_sen += n_sent
        return steps

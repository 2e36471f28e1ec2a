"""Host data model: kana lattice -> flat CSR arrays for the device engine.

Mirrors the reference's lattice construction (decoder/decoder.py:79-135) and vocabulary selection
(decoder/decoder.py:137-151, decoder/decoder_dynamic.py:30-46) but emits flat arrays instead of
Node objects: per sentence a list of frames, frame t holding the nodes ENDING at t in the
reference's order (start ascending, lexicon id ascending, '<unk>' fallback last).
"""
import ctypes as C

import numpy as np

from . import _lib


class Node(object):
    """decoder/decoder.py:17-27 (kept so Decoder.backward_lookup has the reference's shape)."""
    __slots__ = ('start_idx', 'reading_length', 'word_idx', 'word', 'oov_prob', 'char_rnn_step')

    def __init__(self, s, l, idx, word, oov_prob=0.0):
        self.start_idx = s
        self.reading_length = l
        self.word_idx = idx
        self.word = word
        self.oov_prob = oov_prob
        self.char_rnn_step = 0

    def __repr__(self):
        return str((self.start_idx, self.word))


class LatticeBuilder(object):
    """Substring lookup against the reading dictionary with a per-reading cache of in-vocabulary words."""

    def __init__(self, w2i, full_lexicon, full_reading_dict):
        self.w2i = w2i
        self.lexicon = full_lexicon
        self.reading_dict = full_reading_dict
        self._cache = {}
        self.eos = w2i['<eos>']
        self.unk = w2i['<unk>']
        self.max_reading = max((len(k) for k in full_reading_dict), default=0)

    def _words(self, reading):
        """[(word_idx, word)] for a reading: lexicon ids ascending, OOV skipped (decoder.py:97-103)."""
        hit = self._cache.get(reading)
        if hit is None:
            hit = []
            for lex_id in sorted(self.reading_dict[reading]):
                word = self.lexicon[lex_id][0]
                idx = self.w2i.get(word)
                if idx is not None:
                    hit.append((idx, word))
            self._cache[reading] = hit
        return hit

    def build(self, text):
        """Returns frames: list over t=0..T of [(start, word_idx, word)] (decoder.py:79-135)."""
        T = len(text)
        frames = [[] for _ in range(T + 1)]
        frames[0].append((-1, self.eos, '<eos>'))
        rd = self.reading_dict
        for i in range(T):
            # the reference scans every suffix length; readings longer than the longest key cannot match
            for j in range(min(T - i, self.max_reading)):
                sub = text[i:i + j + 1]
                if sub in rd:
                    end = frames[i + j + 1]
                    for idx, word in self._words(sub):
                        end.append((i, idx, word))
                if j == 0 and not frames[i + 1]:
                    frames[i + 1].append((i, self.unk, text[i]))   # decoder.py:129-130
            if self.max_reading == 0 and not frames[i + 1]:
                frames[i + 1].append((i, self.unk, text[i]))
        return frames


def to_backward_lookup(frames):
    """frames -> {frame: [Node]} as Decoder.backward_lookup (decoder.py:86-90)."""
    return {t: [Node(s, 1 if s < 0 else t - s, w, word) for (s, w, word) in fr] for t, fr in enumerate(frames)}


def static_vocab(frames, n_words, samples=0, top_sampling=False, random_sampling=False):
    """decoder/decoder.py:137-151: sorted unique word ids (+ random or top samples)."""
    lv = sorted(set(n[1] for fr in frames for n in fr))
    if samples:
        if random_sampling:
            lv += [int(x) for x in np.random.randint(n_words, size=samples)]
        elif top_sampling:
            lv += list(range(samples))
        lv = sorted(set(lv))
    return lv


def dynamic_vocab(frames, n_words, samples=0, top_sampling=False, random_sampling=False):
    """decoder/decoder_dynamic.py:30-46.  Returns (lv0, news): lv0 is lattice_vocab[0] as the reference
    builds it (duplicates kept); news[i] (i>=1) is sorted(lattice_vocab[i] - lattice_vocab[i-1])."""
    lv0 = sorted(n[1] for n in frames[0])
    if samples:
        if random_sampling:
            lv0 += [int(x) for x in np.random.randint(n_words, size=samples)]
        elif top_sampling:
            lv0 += list(range(samples))
    seen = set(lv0)
    news = [None]
    for i in range(1, len(frames)):
        new = sorted(set(n[1] for n in frames[i]) - seen)
        seen.update(new)
        news.append(new)
    return lv0, news


def dynamic_vocab_final(lv0, news):
    """lattice_vocab as DynamicDecoder leaves it after decode(): every frame's list extended in place
    with the later frames' new words (decoder_dynamic.py:112-127)."""
    out = {}
    cum = sorted(set(lv0))
    for i in range(len(news)):
        if i > 0:
            cum = sorted(set(cum) | set(news[i]))
        base = list(lv0) if i == 0 else list(cum)
        for k in range(i + 1, len(news)):
            base += news[k]
        out[i] = base
    return out


class PackedLattices(object):
    """CSR arrays of a batch of lattices + the ctypes view passed to jlm_decode_batch / jlm_batch_upload."""

    def __init__(self, all_frames, vocab_lists=None, dynamic=None):
        S = len(all_frames)
        self.n_sent = S
        self.frames = all_frames
        self.sent_len = np.array([len(fr) - 1 for fr in all_frames], dtype=np.int32)
        self.frame_ptr_off = np.zeros(S, dtype=np.int64)
        n_fp = int(np.sum(self.sent_len.astype(np.int64) + 2))
        self.frame_ptr = np.zeros(n_fp, dtype=np.int64)
        counts = [[len(f) for f in fr] for fr in all_frames]
        n_nodes = sum(sum(c) for c in counts)
        self.node_start = np.empty(n_nodes, dtype=np.int32)
        self.node_word = np.empty(n_nodes, dtype=np.int32)
        self.node_off = np.zeros(S + 1, dtype=np.int64)
        fp, nn = 0, 0
        for s, fr in enumerate(all_frames):
            self.frame_ptr_off[s] = fp
            self.node_off[s] = nn
            self.frame_ptr[fp] = nn
            for t, f in enumerate(fr):
                k = len(f)
                if k:
                    self.node_start[nn:nn + k] = [n[0] for n in f]
                    self.node_word[nn:nn + k] = [n[1] for n in f]
                nn += k
                self.frame_ptr[fp + t + 1] = nn
            fp += len(fr) + 1
        self.node_off[S] = nn
        self.vocab_ptr = self.vocab_ids = self.vocab_frame_ptr = self.dup_ptr = self.dup_ids = None
        if vocab_lists is not None:
            self.vocab_ptr = np.zeros(S + 1, dtype=np.int64)
            self.vocab_ptr[1:] = np.cumsum([len(v) for v in vocab_lists])
            self.vocab_ids = np.array([x for v in vocab_lists for x in v], dtype=np.int32)
        if dynamic is not None:
            # dynamic[s] = (lv0 with duplicates, news); columns ordered by first appearance
            ids, dups, vfp = [], [], []
            vp, dp = [0], [0]
            for lv0, news in dynamic:
                uniq = sorted(set(lv0))
                extra = list(lv0)
                for x in uniq:
                    extra.remove(x)
                cols = list(uniq)
                ptr = [0, len(cols)]
                for i in range(1, len(news)):
                    cols += news[i]
                    ptr.append(len(cols))
                ids += cols
                dups += extra
                vfp += ptr
                vp.append(len(ids))
                dp.append(len(dups))
            self.vocab_ptr = np.array(vp, dtype=np.int64)
            self.vocab_ids = np.array(ids, dtype=np.int32)
            self.vocab_frame_ptr = np.array(vfp, dtype=np.int32)
            self.dup_ptr = np.array(dp, dtype=np.int64)
            self.dup_ids = np.array(dups if dups else [0], dtype=np.int32)

    def c_struct(self):
        lb = _lib.LatticeBatch()
        lb.n_sent = self.n_sent
        lb.sent_len = _lib.ptr(self.sent_len, C.c_int32)
        lb.frame_ptr_off = _lib.ptr(self.frame_ptr_off, C.c_int64)
        lb.frame_ptr = _lib.ptr(self.frame_ptr, C.c_int64)
        lb.node_start = _lib.ptr(self.node_start, C.c_int32)
        lb.node_word = _lib.ptr(self.node_word, C.c_int32)
        lb.vocab_ptr = _lib.ptr(self.vocab_ptr, C.c_int64)
        lb.vocab_ids = _lib.ptr(self.vocab_ids, C.c_int32)
        lb.vocab_frame_ptr = _lib.ptr(self.vocab_frame_ptr, C.c_int32)
        lb.dup_ptr = _lib.ptr(self.dup_ptr, C.c_int64)
        lb.dup_ids = _lib.ptr(self.dup_ids, C.c_int32)
        return lb

    def node_words(self, s):
        """Flat list of word strings of sentence s, indexed by (absolute node index - node_off[s])."""
        return [n[2] for f in self.frames[s] for n in f]

    def path_words(self, s, node_ids):
        """Word strings of the given absolute node indices of sentence s."""
        words = self._words_cache.get(s) if hasattr(self, '_words_cache') else None
        if words is None:
            if not hasattr(self, '_words_cache'):
                self._words_cache = {}
            words = self._words_cache[s] = self.node_words(s)
        off = int(self.node_off[s])
        return [words[i - off] for i in node_ids]


class NativeLexicon(object):
    """Dictionary side of the lattice builder handed to libjlm_b200 once (jlm_lexicon_create):
    reading -> in-vocabulary (word id, word) pairs in `sorted(lexicon ids)` order (decoder.py:97-103)."""

    def __init__(self, w2i, full_lexicon, full_reading_dict):
        self._lib = _lib.load()
        rptr, wptr, chars, wids, self.entry_words = [0], [0], [], [], []
        for reading, ids in full_reading_dict.items():
            cps = [ord(ch) for ch in reading]
            for lex_id in sorted(ids):
                word = full_lexicon[lex_id][0]
                idx = w2i.get(word)
                if idx is not None:
                    wids.append(idx)
                    self.entry_words.append(word)
            chars += cps
            rptr.append(len(chars))
            wptr.append(len(wids))
        self._arrays = (np.array(rptr, dtype=np.int64), np.array(chars if chars else [0], dtype=np.uint32),
                        np.array(wptr, dtype=np.int64), np.array(wids if wids else [0], dtype=np.int32))
        a = self._arrays
        self.handle = C.c_void_p()
        _lib.check(self._lib.jlm_lexicon_create(len(rptr) - 1, _lib.ptr(a[0], C.c_int64), _lib.ptr(a[1], C.c_uint32),
                                                _lib.ptr(a[2], C.c_int64), _lib.ptr(a[3], C.c_int32),
                                                int(w2i['<eos>']), int(w2i['<unk>']), C.byref(self.handle)))

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None and h.value:
            self._lib.jlm_lexicon_destroy(h)
            self.handle = None


class NativeLattices(object):
    """A batch of lattices built by jlm_lattice_build; same interface as PackedLattices for the decode
    call (c_struct, n_sent, sent_len, path_words), arrays owned by the native object."""

    def __init__(self, lexicon, texts, mode=_lib.DECODE_FULL, extra_ids=None):
        self._lib = lexicon._lib
        self._lexicon = lexicon
        self.texts = list(texts)
        self.n_sent = len(self.texts)
        lens = np.fromiter((len(t) for t in self.texts), dtype=np.int64, count=self.n_sent)
        tptr = np.zeros(self.n_sent + 1, dtype=np.int64)
        np.cumsum(lens, out=tptr[1:])
        joined = ''.join(self.texts)
        cps = np.frombuffer(joined.encode('utf-32-le'), dtype=np.uint32) if joined else np.zeros(1, dtype=np.uint32)
        n_extra = 0
        if extra_ids is not None:
            extra_ids = np.ascontiguousarray(extra_ids, dtype=np.int32).reshape(self.n_sent, -1)
            n_extra = extra_ids.shape[1]
        self.handle = C.c_void_p()
        _lib.check(self._lib.jlm_lattice_build(lexicon.handle, self.n_sent, _lib.ptr(tptr, C.c_int64),
                                               _lib.ptr(cps, C.c_uint32), int(mode), n_extra,
                                               _lib.ptr(extra_ids, C.c_int32) if n_extra else None,
                                               C.byref(self.handle)))
        self._view = _lib.LatticeBatch()
        entry = C.POINTER(C.c_int32)()
        n_nodes = C.c_int64(0)
        _lib.check(self._lib.jlm_lattice_view(self.handle, C.byref(self._view), C.byref(entry), C.byref(n_nodes)))
        self.n_nodes = int(n_nodes.value)
        v = self._view
        self.sent_len = np.ctypeslib.as_array(v.sent_len, shape=(self.n_sent,))
        self.frame_ptr_off = np.ctypeslib.as_array(v.frame_ptr_off, shape=(self.n_sent,))
        n_fp = int(self.sent_len.astype(np.int64).sum()) + 2 * self.n_sent
        self.frame_ptr = np.ctypeslib.as_array(v.frame_ptr, shape=(n_fp,))
        self.node_start = np.ctypeslib.as_array(v.node_start, shape=(self.n_nodes,))
        self.node_word = np.ctypeslib.as_array(v.node_word, shape=(self.n_nodes,))
        self.node_entry = np.ctypeslib.as_array(entry, shape=(self.n_nodes,))
        self.node_off = np.append(self.frame_ptr[self.frame_ptr_off], self.n_nodes).astype(np.int64)
        self.mode = mode

    def __del__(self):
        h = getattr(self, 'handle', None)
        if h is not None and h.value:
            self._lib.jlm_lattice_destroy(h)
            self.handle = None

    def c_struct(self):
        return self._view

    def path_words(self, s, node_ids):
        words, ew = [], self._lexicon.entry_words
        for i in node_ids:
            e = int(self.node_entry[i])
            if e >= 0:
                words.append(ew[e])
            elif e == -1:
                words.append('<eos>')
            else:
                words.append(self.texts[s][int(self.node_start[i])])     # '<unk>' node carries the raw kana
        return words

    def frames_of(self, s):
        """Sentence s as the Python builder's frames (list over t of [(start, word_idx, word)])."""
        fp = self.frame_ptr[int(self.frame_ptr_off[s]):int(self.frame_ptr_off[s]) + int(self.sent_len[s]) + 2]
        out = []
        for t in range(len(fp) - 1):
            ids = range(int(fp[t]), int(fp[t + 1]))
            ws = self.path_words(s, ids)
            out.append([(int(self.node_start[i]), int(self.node_word[i]), w) for i, w in zip(ids, ws)])
        return out

    def vocab_list(self, s):
        v = self._view
        a, b = int(v.vocab_ptr[s]), int(v.vocab_ptr[s + 1])
        return [int(v.vocab_ids[k]) for k in range(a, b)]

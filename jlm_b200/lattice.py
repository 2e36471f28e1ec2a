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

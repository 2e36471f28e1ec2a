"""CharRNNDecoder: word lattice scored by a CHARACTER language model (reference decoder/decoder.py:244-341).

A node's first character is scored from its parent path's distribution; the remaining characters of a
multi-character word take one LM step each, for EVERY candidate path, before the frame is pruned
(_eval_frame, decoder.py:301-320); candidates that spell the same string as an earlier candidate of the
frame are dropped (decoder.py:293-297).  Every LM step runs on the device through the C ABI
(jlm_predict via LSTM_Model.predict_with_context); the candidate bookkeeping, which is per-sentence and
string-keyed, stays on the host as flat arrays.

The reference class does not run as shipped - it reads `self.vocab.words` and indexes `self.w2i` with
characters, while Decoder._load_vocab builds the word-keyed train.data.Vocab (AttributeError at
decoder.py:264; recorded in tests/golden/charrnn_*.json "as_shipped").  This mirror loads what those
lines need, the reference's own CharVocab (train/data.py:28-46): `vocab.words`, w2i = c2i, i2w = i2c.
"""
import math

import numpy as np

from . import lattice
from .decoder import Decoder
from .vocab import CharVocab


class CharRNNDecoder(Decoder):
    char_rnn = True

    def _load_vocab(self):
        self.vocab = CharVocab(self.config['vocab_size'])
        self.i2w = self.vocab.i2c
        self.w2i = self.vocab.c2i

    def _check_oov(self, word):
        # decoder/decoder.py:263-264
        return word not in self.vocab.words

    def _char_check_oov(self, word):
        # decoder/decoder.py:266-267: number of characters of the display string the model does not know
        return sum([c not in self.w2i for c in word.split('/')[0]])

    def _word_length(self, word):
        # decoder/decoder.py:269-273
        return 1 if word in ('<eos>', '<unk>') else len(word)

    # ------------------------------------------------------------------------------------------
    def _reading_nodes(self, reading):
        """[(first character id, display string)] of one reading: lexicon ids ascending, OOV words and words
        with unknown characters skipped, one node per display string, at most 201 strings (decoder.py:95-125)."""
        hit = self._reading_cache.get(reading)
        if hit is None:
            hit, seen = [], set()
            for lex_id in sorted(self.full_reading_dict[reading]):
                word = self.full_lexicon[lex_id][0]
                if self._check_oov(word) or self._char_check_oov(word):
                    continue
                disp = word.split('/')[0]
                if disp in seen or len(seen) > 200:
                    continue
                seen.add(disp)
                hit.append((self.w2i[disp[0]], disp))
            self._reading_cache[reading] = hit
        return hit

    def _build_lattice(self, input, vocab_select=False, samples=0, top_sampling=False, random_sampling=False):
        """decoder/decoder.py:79-135 (char_rnn branch) -> frames[t] = [(start, first char id, display string)]."""
        if not hasattr(self, '_reading_cache'):
            self._reading_cache = {}
            self._max_reading = max((len(k) for k in self.full_reading_dict), default=0)
        T = len(input)
        frames = [[] for _ in range(T + 1)]
        frames[0].append((-1, self.w2i['<eos>'], '<eos>'))
        for i in range(T):
            for j in range(min(T - i, self._max_reading)):
                sub = input[i:i + j + 1]
                if sub in self.full_reading_dict:
                    end = frames[i + j + 1]
                    for cid, disp in self._reading_nodes(sub):
                        end.append((i, cid, disp))
                if j == 0 and not frames[i + 1]:
                    frames[i + 1].append((i, self.w2i['<unk>'], input[i]))
            if self._max_reading == 0 and not frames[i + 1]:
                frames[i + 1].append((i, self.w2i['<unk>'], input[i]))
        if vocab_select:
            # kept for interface parity: CharRNNDecoder.decode never passes it to the model
            self._build_lattice_vocab(frames, samples, top_sampling, random_sampling)
        return frames

    # ------------------------------------------------------------------------------------------
    def _step(self, char_ids, h, c):
        """One batched LM step (decoder.py:202-218): rows in, (probabilities, h, c) out; logs the timers."""
        (pred, _y, t_lstm, t_soft), h2, c2 = self.model.predict_with_context(list(char_ids), h, c, None)
        self.perf_log_lstm.append(t_lstm)
        self.perf_log_softmax.append(t_soft)
        return pred, h2, c2

    def decode(self, input, topN=10, beam_width=10, vocab_select=False, samples=0, top_sampling=False,
               random_sampling=False):
        """decoder/decoder.py:322-341 -> [(neg_log_prob, [display string, ...])][:topN]."""
        frames = self._build_lattice(input, vocab_select=vocab_select, samples=samples, top_sampling=top_sampling,
                                     random_sampling=random_sampling)
        self.backward_lookup = lattice.to_backward_lookup(frames)
        H = self.model.hidden_size
        w2i = self.w2i
        beams = []       # per frame: dict(score [n], h [n,H], c [n,H], probs [n,V], words [n][...], text [n])
        for t, nodes in enumerate(frames):
            if t == 0:
                score = np.zeros(1)
                h = np.zeros((1, H))
                c = np.zeros((1, H))
                words, text, cur = [['<eos>']], ['<eos>'], [w2i['<eos>']]
                wlen = [1]
            else:
                # expand in the reference's order (node, then parent rank); the first spelling of a string wins
                s_l, h_l, c_l, words, text, cur, wlen, seen = [], [], [], [], [], [], [], set()
                for (start, cid, disp) in nodes:
                    par = beams[start]
                    for r in range(len(par['score'])):
                        spelled = par['text'][r] + disp
                        if spelled in seen:
                            continue
                        seen.add(spelled)
                        s_l.append(par['score'][r] + -math.log(par['probs'][r, cid]))     # Path.append_node :43-49
                        h_l.append(par['h'][r])
                        c_l.append(par['c'][r])
                        words.append(par['words'][r] + [disp])
                        text.append(spelled)
                        cur.append(cid)
                        wlen.append(self._word_length(disp))
                score = np.array(s_l)
                h = np.stack(h_l)
                c = np.stack(c_l)
                # the remaining characters of multi-character words, one LM step per character, all
                # candidates of the frame that still have characters left batched together (:301-320)
                k = 1
                active = [a for a in range(len(cur)) if wlen[a] > 1]
                while active:
                    pred, h2, c2 = self._step([cur[a] for a in active], h[active], c[active])
                    h[active], c[active] = h2, c2
                    for row, a in enumerate(active):
                        cur[a] = w2i[words[a][-1][k]]
                        score[a] += -np.log(pred[row, cur[a]])
                    k += 1
                    active = [a for a in active if wlen[a] > k]
            if beam_width is not None:
                keep = np.argsort(score, kind='stable')[:beam_width]      # list.sort is stable (:331-333)
            else:
                keep = np.arange(len(score))
            score, h, c = score[keep], h[keep], c[keep]
            words = [words[a] for a in keep]
            text = [text[a] for a in keep]
            pred, h, c = self._step([cur[a] for a in keep], h, c)        # feeds each survivor's LAST character
            beams.append({'score': score, 'h': h, 'c': c, 'probs': pred, 'words': words, 'text': text})
        self._last_beams = beams
        last = beams[len(input)]
        out = [(float(last['score'][r]), [w for w in last['words'][r] if w != '<eos>']) for r in range(len(last['score']))]
        self.perf_sen += 1
        return out[:topN]

    def decode_batch(self, inputs, topN=10, beam_width=10, vocab_select=False, samples=0, top_sampling=False,
                     random_sampling=False, **kw):
        """decode() for many sentences in lock-step over a device-resident state pool (jlm_pool_*): per frame, ONE
        call scores every (parent state, first character) pair of every sentence, one LM step per remaining
        character position advances every unfinished candidate of every sentence together, and one LM step
        seats the pruned beams.  The string-keyed bookkeeping (de-duplication, pruning) stays on the host; states
        never leave the device.  Same results as decode() sentence by sentence."""
        import time
        from .statepool import StatePool
        inputs = list(inputs)
        S = len(inputs)
        if S == 0:
            return []
        lat = [self._build_lattice(x, vocab_select=vocab_select, samples=samples, top_sampling=top_sampling,
                                   random_sampling=random_sampling) for x in inputs]
        w2i = self.w2i
        # states needed: every candidate takes len(word)-1 steps, every kept path one more
        W = beam_width if beam_width is not None else 1 << 30
        need, bc_hint = 0, []
        for fr in lat:
            bc = [1]
            for t in range(1, len(fr)):
                n_c = sum(bc[n[0]] for n in fr[t])
                need += sum(bc[n[0]] * (self._word_length(n[2]) - 1) for n in fr[t])
                bc.append(min(W, n_c))
            need += sum(bc)
        pool = getattr(self, '_pool', None)
        if pool is None or pool.capacity < need:
            self._pool = pool = StatePool(self.model, max(need, 1024))
        pool.reset()
        # Per (sentence, frame) the kept paths are flat arrays: score, state slot, spelled text, and a back-pointer
        # (start frame, rank there, display string) - word lists are only materialised for the final n-best.
        beams = [[] for _ in range(S)]
        Tmax = max(len(x) for x in inputs)
        eos = w2i['<eos>']
        for t in range(Tmax + 1):
            act = [s for s in range(S) if len(inputs[s]) >= t]
            # ---- expansion: flat candidate arrays over all active sentences, sentence by sentence ----
            off = [0]
            if t == 0:
                n_act = len(act)
                score = np.zeros(n_act)
                slot = np.full(n_act, -1, dtype=np.int64)
                cur = np.full(n_act, eos, dtype=np.int64)
                wlen = np.ones(n_act, dtype=np.int64)
                text = ['<eos>'] * n_act
                disp = ['<eos>'] * n_act
                bp_f = [-1] * n_act
                bp_r = [0] * n_act
                off = list(range(n_act + 1))
            else:
                base_l, slot_l, cur_l, wlen_l, text, disp, bp_f, bp_r = [], [], [], [], [], [], [], []
                for s in act:
                    seen = set()
                    frames_s = beams[s]
                    for (start, cid, dsp) in lat[s][t]:
                        par = frames_s[start]
                        wl = self._word_length(dsp)
                        ptext, pscore, pslot = par['text'], par['score'], par['slot']
                        for r in range(len(ptext)):
                            spelled = ptext[r] + dsp
                            if spelled in seen:                      # first spelling wins (decoder.py:293-297)
                                continue
                            seen.add(spelled)
                            text.append(spelled)
                            base_l.append(pscore[r])
                            slot_l.append(pslot[r])
                            bp_r.append(r)
                            bp_f.append(start)
                            cur_l.append(cid)
                            wlen_l.append(wl)
                            disp.append(dsp)
                    off.append(len(text))
                slot = np.asarray(slot_l, dtype=np.int64)
                cur = np.asarray(cur_l, dtype=np.int64)
                wlen = np.asarray(wlen_l, dtype=np.int64)
                t0 = time.time()
                score = np.asarray(base_l, dtype=np.float64) + pool.nll(slot, cur)   # Path.append_node (decoder.py:43-49)
                self.perf_log_softmax.append(time.time() - t0)
                # remaining characters of multi-character words: one LM step per position, all sentences together
                k = 1
                active = np.nonzero(wlen > 1)[0]
                while len(active):
                    t0 = time.time()
                    new = pool.step(slot[active], cur[active])
                    self.perf_log_lstm.append(time.time() - t0)
                    nxt = np.fromiter((w2i[disp[a][k]] for a in active), dtype=np.int64, count=len(active))
                    t0 = time.time()
                    score[active] += pool.nll(new, nxt)
                    self.perf_log_softmax.append(time.time() - t0)
                    slot[active] = new
                    cur[active] = nxt
                    k += 1
                    active = active[wlen[active] > k]
            # ---- prune per sentence (stable sort, decoder.py:331-333), then seat the survivors ----
            keeps = []
            for i in range(len(act)):
                lo, hi = off[i], off[i + 1]
                if beam_width is not None:
                    kp = np.argsort(score[lo:hi], kind='stable')[:beam_width] + lo
                else:
                    kp = np.arange(lo, hi)                            # no sort either
                keeps.append(kp)
            allk = np.concatenate(keeps) if keeps else np.zeros(0, dtype=np.int64)
            t0 = time.time()
            new = pool.step(slot[allk], cur[allk])
            self.perf_log_lstm.append(time.time() - t0)
            o = 0
            for i, s in enumerate(act):
                kp = keeps[i]
                n = len(kp)
                kl = kp.tolist()
                beams[s].append({'score': score[kp], 'slot': new[o:o + n], 'text': [text[a] for a in kl],
                                 'bp_f': [bp_f[a] for a in kl], 'bp_r': [bp_r[a] for a in kl],
                                 'disp': [disp[a] for a in kl]})
                o += n
        out = []
        for s in range(S):
            T = len(inputs[s])
            last = beams[s][T]
            res = []
            for r in range(len(last['score'])):
                words, f, k = [], T, r
                while f >= 0:
                    fr = beams[s][f]
                    words.append(fr['disp'][k])
                    f, k = fr['bp_f'][k], fr['bp_r'][k]
                words.reverse()
                res.append((float(last['score'][r]), [w for w in words if w != '<eos>']))
            out.append(res[:topN])
        self._last_batch_beams = beams
        self.perf_sen += S
        return out

    def decode_stream(self, *a, **kw):
        raise NotImplementedError('decode_stream is the full-softmax word decoder\'s lock-step path')

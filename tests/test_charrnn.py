"""Char-RNN decoder (reference decoder/decoder.py:244-341).

CPU: the oracle restatement against fixtures written by the reference class (tests/golden/charrnn_*.json,
generator tests/golden/make_golden_charrnn.py - the class needs the CharVocab attributes supplied to run at
all; the fixture records the as-shipped AttributeError).  GPU: jlm_b200.CharRNNDecoder, whose LM steps go
through the C ABI, against the same fixtures and against the oracle on fresh sentences."""
import json
import os

import numpy as np
import pytest

from jlm_b200 import synth
from oracle import jlm_oracle as O
from tests.golden.cases import CHAR_CASES
from tests.helpers import GOLDEN

TOL = 1e-9


def _case(name):
    case = CHAR_CASES[name]
    meta = json.load(open(os.path.join(GOLDEN, name + '.json')))
    return case, meta


def _experiment(case, root):
    cfg, weights, lexicon, reading_dict = synth.make_char_experiment(
        str(root), 1, case['vocab_size'], case['hidden_size'], case['embed_size'], seed=case['seed'])
    return cfg, weights, lexicon, reading_dict


def _check_weights(meta, weights):
    chk = float(sum(float(np.sum(np.asarray(v, dtype=np.float64))) for k, v in sorted(weights.items())
                    if not isinstance(v, list)))
    assert chk == meta['weights_checksum'], 'synthetic generator drifted from the golden fixture'


@pytest.mark.parametrize('name', sorted(CHAR_CASES))
def test_oracle_charrnn_matches_reference(name, tmp_path):
    case, meta = _case(name)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    _check_weights(meta, weights)
    # the unmodified class cannot run (decoder.py:264); the fixture pins that too
    assert meta['as_shipped'].startswith("AttributeError: 'Vocab' object has no attribute 'words'")
    words, c2i = O.make_char_vocab(lexicon, cfg['vocab_size'])
    assert len(c2i) == meta['n_chars'] == weights['LM'].shape[0]
    model = O.OracleModel(cfg, weights)
    kw = case['decode_kwargs']
    sentences = synth.make_char_sentences(lexicon, case['n_sent'], min_len=case['min_len'], seed=case['seed'] + 1,
                                          vocab_size=case['vocab_size'])
    assert sentences == meta['sentences']
    dropped = 0
    for sent, g in zip(sentences, meta['decode']):
        frames = O.build_lattice_char(sent, words, c2i, lexicon, reading_dict)
        assert {str(t): [[n[0], n[1], n[2]] for n in fr] for t, fr in enumerate(frames) if fr} == g['lattice']
        trace = []
        res = O.decode_charrnn(model, frames, c2i, kw['topN'], kw['beam_width'], trace=trace)
        assert [ws for _, ws in res] == [ws for _, ws in g['nbest']]
        np.testing.assert_allclose([s for s, _ in res], [s for s, _ in g['nbest']], rtol=0, atol=TOL)
        # candidates kept after string de-duplication, and the rows of every LM call, frame by frame
        assert [t['n_candidates'] for t in trace] == g['n_candidates']
        rows = []
        for t in trace:
            rows += t['eval_rows'] + [len(t['paths'])]
        assert rows == g['predict_rows']
        for t, gp in zip(trace, g['pruned']):
            assert [p[1] for p in t['paths']] == [p[1] for p in gp]
            np.testing.assert_allclose([p[0] for p in t['paths']], [p[0] for p in gp], rtol=0, atol=TOL)
        raw = [1] + [sum(len(trace[n[0]]['paths']) for n in frames[t]) for t in range(1, len(frames))]
        dropped += sum(raw) - sum(g['n_candidates'])
    if name == 'charrnn_beam20':
        assert dropped > 0, 'fixture no longer exercises string de-duplication'


def test_char_vocab_mirror_matches_oracle():
    from jlm_b200.vocab import CharVocab
    lexicon, _ = synth.make_char_lexicon(500, seed=5)
    v = CharVocab(500, lexicon=lexicon)
    words, c2i = O.make_char_vocab(lexicon, 500)
    assert v.c2i == c2i == synth.char_vocab(lexicon, 500) and v.words == words
    assert v.i2c[0] == '<unk>' and v.i2c[1] == '<eos>' and len(v) == len(c2i)


# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CHAR_CASES))
def test_gpu_charrnn_matches_reference(name, tmp_path):
    import jlm_b200
    from jlm_b200 import config
    case, meta = _case(name)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    _check_weights(meta, weights)
    config.set_root(str(tmp_path))
    with pytest.raises(ValueError):
        jlm_b200.Decoder(1)                       # a char_rnn experiment needs the char decoder
    dec = jlm_b200.CharRNNDecoder(1)
    assert dec._check_oov('no such word') and not dec._check_oov(lexicon[1][0])
    for sent, g in zip(meta['sentences'], meta['decode']):
        n0 = len(dec.perf_log_lstm)
        res = dec.decode(sent, **case['decode_kwargs'])
        assert {str(t): [[n.start_idx, n.word_idx, n.word] for n in fr]
                for t, fr in dec.backward_lookup.items() if fr} == g['lattice']
        assert [ws for _, ws in res] == [ws for _, ws in g['nbest']]            # same strings, same order
        np.testing.assert_allclose([s for s, _ in res], [s for s, _ in g['nbest']], rtol=0, atol=2e-5)
        assert len(dec.perf_log_lstm) - n0 == len(g['predict_rows'])           # one timer entry per LM call
        for t, gp in enumerate(g['pruned']):
            b = dec._last_beams[t]
            assert [[w for w in ws] for ws in b['words']] == [[n[1] for n in p[1]] for p in gp]
            np.testing.assert_allclose(b['score'], [p[0] for p in gp], rtol=0, atol=2e-5)
    assert dec.perf_sen == len(meta['sentences'])


@pytest.mark.gpu
def test_gpu_charrnn_matches_oracle_on_fresh_sentences(tmp_path):
    import jlm_b200
    from jlm_b200 import config
    case = dict(CHAR_CASES['charrnn_small'], seed=11, hidden_size=96, embed_size=48)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    config.set_root(str(tmp_path))
    dec = jlm_b200.CharRNNDecoder(1)
    words, c2i = O.make_char_vocab(lexicon, cfg['vocab_size'])
    model = O.OracleModel(cfg, weights)
    sents = synth.make_char_sentences(lexicon, 12, min_len=14, seed=99, vocab_size=case['vocab_size']) + ['ヰ', 'ヰヱ']
    short = [x[:6] for x in sents[:4]] + sents[-2:]
    for beam, batch in ((1, sents), (8, sents), (None, short)):       # beam_width=None: no pruning (decoder.py:331)
        got = dec.decode_batch(batch, topN=5, beam_width=beam)
        for sent, res in zip(batch, got):
            frames = O.build_lattice_char(sent, words, c2i, lexicon, reading_dict)
            want = O.decode_charrnn(model, frames, c2i, 5, beam)
            assert [ws for _, ws in res] == [ws for _, ws in want]
            np.testing.assert_allclose([s for s, _ in res], [s for s, _ in want], rtol=0, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['tied', 'untied', 'dsoftmax_star'])
def test_gpu_state_pool_matches_predict_with_context(mode, tmp_path):
    """jlm_pool_step / jlm_pool_nll against LSTM_Model.predict_with_context (model.py:195-198) on the same rows:
    states, and -log p of arbitrary (state, word) pairs; chained steps, zero-state rows, reset."""
    import jlm_b200
    from jlm_b200 import config
    from jlm_b200.statepool import StatePool
    synth.make_experiment(str(tmp_path), 1, 300, 48, 24, mode, seed=9, write_lexicon=False)
    config.set_root(str(tmp_path))
    m = jlm_b200.LSTM_Model(1)
    pool = StatePool(m, 64)
    rng = np.random.default_rng(0)
    idx1 = rng.integers(0, 300, size=5)
    (p1, _, _, _), h1, c1 = m.predict_with_context(idx1.tolist(), np.zeros((5, 48)), np.zeros((5, 48)))
    s1 = pool.step([-1] * 5, idx1)
    assert s1.tolist() == [0, 1, 2, 3, 4]
    hp, cp = pool.state(0, 5)
    np.testing.assert_allclose(hp, h1, rtol=0, atol=1e-12)
    np.testing.assert_allclose(cp, c1, rtol=0, atol=1e-12)
    cols = rng.integers(0, 300, size=5)
    np.testing.assert_allclose(pool.nll(s1, cols), -np.log(p1[np.arange(5), cols]), rtol=0, atol=1e-9)
    # second step from a permutation of the first states plus one fresh zero-state row
    src = [3, 0, 0, -1]
    idx2 = rng.integers(0, 300, size=4)
    hin = np.stack([h1[3], h1[0], h1[0], np.zeros(48)])
    cin = np.stack([c1[3], c1[0], c1[0], np.zeros(48)])
    (p2, _, _, _), h2, c2 = m.predict_with_context(idx2.tolist(), hin, cin)
    s2 = pool.step(src, idx2)
    assert s2.tolist() == [5, 6, 7, 8]
    hp, cp = pool.state(5, 4)
    np.testing.assert_allclose(hp, h2, rtol=0, atol=1e-12)
    pairs_s = [5, 5, 8, 2, 7]
    pairs_c = [0, 299, 17, 4, 4]
    want = [-np.log(p2[0, 0]), -np.log(p2[0, 299]), -np.log(p2[3, 17]), -np.log(p1[2, 4]), -np.log(p2[2, 4])]
    np.testing.assert_allclose(pool.nll(pairs_s, pairs_c), want, rtol=0, atol=1e-9)
    with pytest.raises(jlm_b200._lib.JlmError):
        pool.step([9], [1])                       # slot 9 holds no state yet
    with pytest.raises(jlm_b200._lib.JlmError):
        pool.step([-1] * 100, [1] * 100)          # over capacity
    pool.reset()
    assert pool.step([-1], [int(idx1[0])]).tolist() == [0]


@pytest.mark.gpu
def test_gpu_charrnn_lockstep_batch_equals_per_sentence(tmp_path):
    import jlm_b200
    from jlm_b200 import config
    case = dict(CHAR_CASES['charrnn_small'], seed=5)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    config.set_root(str(tmp_path))
    dec = jlm_b200.CharRNNDecoder(1)
    sents = synth.make_char_sentences(lexicon, 40, min_len=4, seed=123, vocab_size=case['vocab_size'])
    sents = [s[:3 + (i % 14)] for i, s in enumerate(sents)] + ['', 'ヰ']
    for beam, topn in ((4, 4), (12, 3)):
        n0 = dec.perf_sen
        got = dec.decode_batch(sents, topN=topn, beam_width=beam)
        assert dec.perf_sen - n0 == len(sents)
        for sent, g in zip(sents, got):
            want = dec.decode(sent, topN=topn, beam_width=beam)
            assert [ws for _, ws in g] == [ws for _, ws in want], sent
            np.testing.assert_allclose([s for s, _ in g], [s for s, _ in want], rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_gpu_eval_harness_runs_a_char_rnn_experiment(tmp_path):
    """eval.py:35-44 picks CharVocab + CharRNNDecoder from config['char_rnn']; the harness must drive it (lock-step
    batch and per-pair) and count hits on display strings.  A character LM that copies nothing still has to find
    every target among the lattice paths when the beam is wide open on short inputs."""
    import jlm_b200
    from jlm_b200 import config, eval as jeval
    case = dict(CHAR_CASES['charrnn_small'], seed=2)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    lines = synth.make_test_corpus(lexicon[:cfg['vocab_size'] - 1], 24, seed=4, min_words=1, max_words=2)
    synth.write_test_corpus(str(tmp_path), lines)
    config.set_root(str(tmp_path))
    logs = []
    for extra in ([], ['--no-batch']):
        args = jeval.build_parser().parse_args(['-e', '1', '-es', '12', '-b', '64', '--log_dir', str(tmp_path / 'eval')] + extra)
        ev = jeval.Evaluator(args)
        assert isinstance(ev.decoder, jlm_b200.CharRNNDecoder)
        best, nbest, miss, n = ev.evaluate()
        assert n == 12 and best + nbest + miss == 12
        logs.append(open(ev.log_path, encoding='utf-8').read().split('best_hit')[0])
    assert logs[0] == logs[1]                     # lock-step batch and per-pair decoding write the same log


@pytest.mark.parametrize('name', sorted(CHAR_CASES))
def test_mirror_char_lattice_matches_reference_on_cpu(name, tmp_path):
    """CharRNNDecoder._build_lattice (host code, no device needed) against the lattices the reference class built:
    display strings, first-character ids, per-reading de-duplication, '<unk>' fallback, order."""
    import pickle
    from jlm_b200 import config
    from jlm_b200.decoder_charrnn import CharRNNDecoder
    from jlm_b200.vocab import CharVocab
    case, meta = _case(name)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    config.set_root(str(tmp_path))
    dec = object.__new__(CharRNNDecoder)             # no LSTM_Model: only the host-side lattice code is exercised
    dec.config = cfg
    dec.full_lexicon, dec.full_reading_dict = lexicon, reading_dict
    dec.lattice_vocab = None
    dec._load_vocab()
    assert isinstance(dec.vocab, CharVocab) and len(dec.w2i) == meta['n_chars']
    for g in meta['decode']:
        frames = dec._build_lattice(g['input'], vocab_select=case['decode_kwargs'].get('vocab_select', False))
        assert {str(t): [[n[0], n[1], n[2]] for n in fr] for t, fr in enumerate(frames) if fr} == g['lattice']
    # the 201-strings-per-reading cap (decoder.py:119): 260 homophones with distinct spellings
    many = [('<eos>', 10)] + [('{}{}/ア/P'.format(chr(0x4E00 + k // 40), chr(0x4E00 + k % 40)), 5) for k in range(260)]
    dec2 = object.__new__(CharRNNDecoder)
    dec2.config = dict(cfg, vocab_size=len(many) + 1)
    dec2.full_lexicon, dec2.full_reading_dict = many, {'ア': list(range(1, 261))}
    dec2.lattice_vocab = None
    pickle.dump(many, open(str(tmp_path / 'data' / 'lexicon.pkl'), 'wb'))
    dec2._load_vocab()
    fr = dec2._build_lattice('ア')
    assert len(fr[1]) == 201
    words, c2i = O.make_char_vocab(many, len(many) + 1)
    assert [[n[0], n[1], n[2]] for n in fr[1]] == [[n[0], n[1], n[2]] for n in
                                                  O.build_lattice_char('ア', words, c2i, many, {'ア': list(range(1, 261))})[1]]


class _CpuModel(object):
    """LSTM_Model surface over the numpy oracle: lets the decoder's HOST logic run without a device."""

    def __init__(self, cfg, weights):
        self.om = O.OracleModel(cfg, weights)
        self.hidden_size = self.om.hidden_size

    def predict_with_context(self, index, hidden, cell, vocab=None):
        pred, y, h, c = self.om.predict(list(index), np.asarray(hidden), np.asarray(cell), vocab)
        return (pred, y, 0.0, 0.0), h, c


class _CpuPool(object):
    """StatePool surface (step / nll / reset / capacity) over the numpy oracle."""

    def __init__(self, model):
        self.m, self.capacity = model, 1 << 30
        self.reset()

    def reset(self):
        self.h, self.c, self.p = [], [], []
        self.used = 0

    def step(self, src, index):
        H = self.m.hidden_size
        hin = np.stack([self.h[s] if s >= 0 else np.zeros(H) for s in src])
        cin = np.stack([self.c[s] if s >= 0 else np.zeros(H) for s in src])
        (pred, _, _, _), h, c = self.m.predict_with_context([int(i) for i in index], hin, cin)
        first = len(self.h)
        for k in range(len(src)):
            self.h.append(h[k]); self.c.append(c[k]); self.p.append(pred[k])
        self.used = len(self.h)
        return np.arange(first, first + len(src), dtype=np.int64)

    def nll(self, slots, cols):
        return np.array([-np.log(self.p[int(s)][int(c)]) for s, c in zip(slots, cols)])


@pytest.mark.parametrize('name', sorted(CHAR_CASES))
def test_mirror_char_decoder_host_logic_on_cpu(name, tmp_path):
    """decode() (per-call shape) and the lock-step decode_batch() of jlm_b200.CharRNNDecoder with the LM replaced by
    the numpy oracle: expansion order, string de-duplication, multi-step word evaluation, stable pruning and the
    slot bookkeeping are host code and must reproduce the reference fixtures without a GPU."""
    from jlm_b200 import config
    from jlm_b200.decoder_charrnn import CharRNNDecoder
    case, meta = _case(name)
    cfg, weights, lexicon, reading_dict = _experiment(case, tmp_path)
    config.set_root(str(tmp_path))
    dec = object.__new__(CharRNNDecoder)
    dec.config = cfg
    dec.full_lexicon, dec.full_reading_dict = lexicon, reading_dict
    dec.lattice_vocab = None
    dec.perf_sen, dec.perf_log_lstm, dec.perf_log_softmax = 0, [], []
    dec._load_vocab()
    dec.model = _CpuModel(cfg, weights)
    dec._pool = _CpuPool(dec.model)
    kw = case['decode_kwargs']
    sents = meta['sentences']
    batch = dec.decode_batch(sents, **kw)
    for sent, g, got_b in zip(sents, meta['decode'], batch):
        got = dec.decode(sent, **kw)
        for res in (got, got_b):
            assert [ws for _, ws in res] == [ws for _, ws in g['nbest']]
            np.testing.assert_allclose([s for s, _ in res], [s for s, _ in g['nbest']], rtol=0, atol=1e-9)
        for t, gp in enumerate(g['pruned']):
            assert dec._last_beams[t]['words'] == [[n[1] for n in p[1]] for p in gp]
    assert dec.perf_sen == 2 * len(sents)

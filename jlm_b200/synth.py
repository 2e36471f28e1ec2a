"""Seeded synthetic JLM experiments (lexicon, reading dict, weights, sentences).

The reference ships no data, weights or lexicons (its .gitignore excludes ``data/``, ``*.pkl``),
so every parity test and benchmark runs on a synthetic experiment directory with exactly the
on-disk layout the reference reads:

    <root>/data/lexicon.pkl                       list[(word, freq)]   (reference data.py:33,44)
    <root>/data/reading_dict.pkl                  {reading: [lexicon idx]} (reference data.py:66-86)
    <root>/train/experiments/<id>/config.json     keys read at decoder/model.py:39-56, decoder.py:56,72
    <root>/train/experiments/<id>/weights/lstm_weights.pkl   {name: float32 ndarray}
                                                  (schema: train/weights.py:30-58)

Generators follow SURVEY.md section 8(d): 80 kana, reading pool of V//3 strings, Zipf(0.6)
reading assignment, weights scaled so the LM is peaked (std of logits ~ 3).
All randomness comes from ``np.random.default_rng(seed)``.
"""
import json
import os
import pickle

import numpy as np

KANA = [chr(0x30A1 + k) for k in range(80)]
_LEN_P = np.array([.01, .19, .35, .28, .12, .05])

# projection modes (decoder/model.py:141-193)
MODE_UNTIED = "untied"          # share_embedding=False: h.UM + b2
MODE_TIED = "tied"              # share_embedding=True : (h.PM).LM^T + b2
MODE_DSOFTMAX = "dsoftmax"      # D_softmax=True      : block-diagonal LM
MODE_DSOFTMAX_STAR = "dsoftmax_star"  # V_table=True  : LM_i . VT_i

MODES = (MODE_UNTIED, MODE_TIED, MODE_DSOFTMAX, MODE_DSOFTMAX_STAR)


def make_lexicon(vocab_size, seed=0, extra_oov=1000):
    """Returns (lexicon, reading_dict) in the reference pickle formats."""
    rng = np.random.default_rng(seed)
    n_pool = max(vocab_size // 3, 8)
    lens = rng.choice(np.arange(1, 7), size=n_pool, p=_LEN_P)
    pool, seen = [], set()
    for L in lens:
        r = ''.join(KANA[k] for k in rng.integers(0, len(KANA), size=int(L)))
        if r not in seen:
            seen.add(r)
            pool.append(r)
    n_words = vocab_size - 2 + extra_oov
    # Zipf(0.6) over the pool
    w = 1.0 / np.power(np.arange(1, len(pool) + 1), 0.6)
    w /= w.sum()
    ridx = rng.choice(len(pool), size=n_words, p=w)
    lexicon = [('<eos>', 10 ** 8)]
    reading_dict = {}
    for n in range(n_words):
        reading = pool[int(ridx[n])]
        lexicon.append(('w{}/{}/P'.format(n, reading), 10 ** 8 // (n + 1)))
        reading_dict.setdefault(reading, []).append(n + 1)  # index into the lexicon list
    return lexicon, reading_dict


def make_sentences(lexicon, n_sent, min_len=20, seed=1, vocab_size=None):
    """Kana strings: concatenated readings of words sampled proportionally to frequency
    until >= min_len kana (mirrors how eval.py:152 builds readings from corpus lines)."""
    rng = np.random.default_rng(seed)
    words = lexicon[1:(vocab_size - 1 if vocab_size else len(lexicon))]
    freq = np.array([f for _, f in words], dtype=np.float64)
    freq /= freq.sum()
    cdf = np.cumsum(freq)
    out = []
    for _ in range(n_sent):
        s = ''
        while len(s) < min_len:
            k = int(np.searchsorted(cdf, rng.random()))
            k = min(k, len(words) - 1)
            s += words[k][0].split('/')[1]
        out.append(s)
    return out


def default_segments(mode, vocab_size, embed_size):
    """Three segments with widths E, E/2, E/4 at 24%/60%/100% of V (README config block shape)."""
    a = int(vocab_size * 0.24)
    b = int(vocab_size * 0.60)
    return [[embed_size, 0, a], [embed_size // 2, a, b], [embed_size // 4, b, None]]


def make_config(vocab_size, hidden_size, embed_size, mode, segments=None, self_norm=False):
    cfg = {
        'vocab_size': vocab_size,
        'hidden_size': hidden_size,
        'embed_size': embed_size,
        'share_embedding': mode != MODE_UNTIED,
        'D_softmax': mode == MODE_DSOFTMAX,
        'V_table': mode == MODE_DSOFTMAX_STAR,
        'embedding_seg': segments if segments is not None else default_segments(mode, vocab_size, embed_size),
        'self_norm': self_norm,
        'char_rnn': False,
    }
    return cfg


def _seg_bounds(cfg):
    V = cfg['vocab_size']
    return [(int(sz), int(s), V if e is None else int(e)) for sz, s, e in cfg['embedding_seg']]


def input_embed_size(cfg):
    """Width of the LSTM input embedding (decoder/model.py:43,49,56-71)."""
    if cfg['D_softmax']:
        return sum(sz for sz, _, _ in _seg_bounds(cfg))
    if cfg['V_table']:
        return _seg_bounds(cfg)[0][0]
    return cfg['embed_size']


def make_weights(cfg, seed=0, logit_std=3.0):
    """float32 weight dict with the reference's key names (train/weights.py:30-58)."""
    rng = np.random.default_rng(seed + 1000)
    V, H = cfg['vocab_size'], cfg['hidden_size']
    E_in = input_embed_size(cfg)
    f32 = np.float32
    w = {}
    a = 1.0 / np.sqrt(H)
    for g in 'ifog':
        w['HM' + g] = rng.uniform(-a, a, size=(H, H)).astype(f32)
        w['IM' + g] = rng.uniform(-a, a, size=(E_in, H)).astype(f32)
        w['b' + g] = rng.normal(0, 0.05, size=(H,)).astype(f32)
    w['b2'] = rng.normal(0, 0.5, size=(V,)).astype(f32)
    segs = _seg_bounds(cfg)
    if not cfg['share_embedding']:
        w['LM'] = rng.normal(0, 0.5, size=(V, E_in)).astype(f32)
        w['UM'] = rng.normal(0, 1.0, size=(H, V)).astype(f32)
        scale_key = 'UM'
    elif cfg['D_softmax']:
        w['LM'] = [rng.normal(0, 0.5, size=(e - s, sz)).astype(f32) for sz, s, e in segs]
        w['PM'] = rng.normal(0, 1.0, size=(H, E_in)).astype(f32)
        scale_key = 'PM'
    elif cfg['V_table']:
        for i, (sz, s, e) in enumerate(segs):
            w['LM{}'.format(i)] = rng.normal(0, 0.5, size=(e - s, sz)).astype(f32)
            if i:
                w['VT{}'.format(i)] = rng.normal(0, 1.0 / np.sqrt(sz), size=(sz, segs[0][0])).astype(f32)
        w['PM'] = rng.normal(0, 1.0, size=(H, segs[0][0])).astype(f32)
        scale_key = 'PM'
    else:
        w['LM'] = rng.normal(0, 0.5, size=(V, E_in)).astype(f32)
        w['PM'] = rng.normal(0, 1.0, size=(H, E_in)).astype(f32)
        scale_key = 'PM'
    # calibrate the projection so the LM is peaked: std(y - b2) ~= logit_std after one step from <eos>
    y = _probe_logits(cfg, w)
    s = float(np.std(y))
    if s > 0:
        w[scale_key] = (w[scale_key] * (logit_std / s)).astype(f32)
    return w


def _input_table(cfg, w):
    segs = _seg_bounds(cfg)
    if cfg['D_softmax']:
        E_in = sum(sz for sz, _, _ in segs)
        LM = np.zeros((cfg['vocab_size'], E_in))
        c = 0
        for (sz, s, e), blk in zip(segs, w['LM']):
            LM[s:e, c:c + sz] = blk
            c += sz
        return LM
    if cfg['V_table']:
        parts = [w['LM0']] + [np.dot(w['LM{}'.format(i)], w['VT{}'.format(i)]) for i in range(1, len(segs))]
        return np.concatenate(parts, axis=0)
    return w['LM']


def _probe_logits(cfg, w, n=8):
    """Bias-free logits for a few rows one LSTM step away from the zero state (calibration only)."""
    H = cfg['hidden_size']
    LM = _input_table(cfg, w)
    idx = np.arange(1, 1 + n)
    e = LM[idx].astype(np.float64)
    sig = lambda x: 1 / (np.exp(-x) + 1)
    gi = sig(e @ w['IMi'] + w['bi'])
    go = sig(e @ w['IMo'] + w['bo'])
    gg = np.tanh(e @ w['IMg'] + w['bg'])
    h = np.tanh(gg * gi) * go
    segs = _seg_bounds(cfg)
    if not cfg['share_embedding']:
        return h @ w['UM']
    t = h @ w['PM']
    if cfg['D_softmax']:
        ys, c = [], 0
        for (sz, s, e_), blk in zip(segs, w['LM']):
            ys.append(t[:, c:c + sz] @ blk.T)
            c += sz
        return np.concatenate(ys, axis=1)
    if cfg['V_table']:
        ys = [t @ w['LM0'].T]
        for i in range(1, len(segs)):
            ys.append((t @ w['VT{}'.format(i)].T) @ w['LM{}'.format(i)].T)
        return np.concatenate(ys, axis=1)
    return t @ w['LM'].T


def write_experiment(root, experiment_id, cfg, weights, lexicon=None, reading_dict=None):
    """Writes the directory layout the reference (config.py:15-19) and this package both read."""
    data = os.path.join(root, 'data')
    exp = os.path.join(root, 'train', 'experiments', str(experiment_id))
    os.makedirs(data, exist_ok=True)
    os.makedirs(os.path.join(exp, 'weights'), exist_ok=True)
    if lexicon is not None:
        with open(os.path.join(data, 'lexicon.pkl'), 'wb') as f:
            pickle.dump(lexicon, f)
        with open(os.path.join(data, 'reading_dict.pkl'), 'wb') as f:
            pickle.dump(reading_dict, f)
    with open(os.path.join(exp, 'config.json'), 'wt') as f:
        json.dump(cfg, f)
    with open(os.path.join(exp, 'weights', 'lstm_weights.pkl'), 'wb') as f:
        pickle.dump(weights, f)
    return exp


def make_experiment(root, experiment_id, vocab_size, hidden_size, embed_size, mode,
                    segments=None, self_norm=False, seed=0, write_lexicon=True):
    """One call: lexicon + reading dict + config + weights on disk. Returns (cfg, weights, lexicon, reading_dict)."""
    cfg = make_config(vocab_size, hidden_size, embed_size, mode, segments, self_norm)
    weights = make_weights(cfg, seed=seed)
    lexicon = reading_dict = None
    if write_lexicon:
        lexicon, reading_dict = make_lexicon(vocab_size, seed=seed)
    write_experiment(root, experiment_id, cfg, weights, lexicon, reading_dict)
    return cfg, weights, lexicon, reading_dict


def compress_weights(weights, bits=8, seed=0):
    """Stand-in for train/comp.py:20-80 with the same outputs: per tensor a (code uint8, codebook
    [2**bits, 1] float32) pair and the decoded tensor np.take(codebook, code).  Centroids are quantiles
    refined by two Lloyd iterations instead of sklearn's full k-means (the formats, not the clustering
    quality, are what the inference path depends on)."""
    n = 2 ** bits
    dump, decoded = {}, {}
    for key, value in weights.items():
        if isinstance(value, list):
            raise ValueError('comp.py cannot compress the block list of a D-softmax dump')
        flat = np.asarray(value, dtype=np.float32).reshape(-1).astype(np.float64)
        cent = np.unique(np.quantile(flat, (np.arange(n) + 0.5) / n))
        for _ in range(2):
            edges = (cent[1:] + cent[:-1]) / 2
            code = np.searchsorted(edges, flat)
            sums = np.bincount(code, weights=flat, minlength=len(cent))
            cnt = np.bincount(code, minlength=len(cent))
            cent = np.where(cnt > 0, sums / np.maximum(cnt, 1), cent)
            cent = np.sort(cent)
        edges = (cent[1:] + cent[:-1]) / 2
        code = np.searchsorted(edges, flat).astype(np.uint8).reshape(np.asarray(value).shape)
        codebook = cent.astype(np.float32).reshape(-1, 1)
        dump[key] = (code, codebook)
        decoded[key] = np.take(codebook, code)
    return dump, decoded


def write_compressed(root, experiment_id, weights, bits=8):
    """Writes lstm_weights_comp_N.pkl and comp_N/lstm_weights_comp_dump.pkl (train/comp.py:52-80 layout)."""
    wdir = os.path.join(root, 'train', 'experiments', str(experiment_id), 'weights')
    os.makedirs(os.path.join(wdir, 'comp_{}'.format(bits)), exist_ok=True)
    dump, decoded = compress_weights(weights, bits)
    with open(os.path.join(wdir, 'lstm_weights_comp_{}.pkl'.format(bits)), 'wb') as f:
        pickle.dump(decoded, f)
    with open(os.path.join(wdir, 'comp_{}'.format(bits), 'lstm_weights_comp_dump.pkl'), 'wb') as f:
        pickle.dump(dump, f)
    return dump, decoded


def make_test_corpus(lexicon, n_lines, seed=2, min_words=3, max_words=8):
    """Lines of data/test.txt as eval.py reads them (decoder/eval.py:125-166): space separated
    'display/reading/POS' tokens.  Words are drawn by frequency from the WHOLE lexicon, so some lines
    contain out-of-vocabulary words and exercise the reference's has_oov filter."""
    rng = np.random.default_rng(seed)
    words = lexicon[1:]
    freq = np.array([f for _, f in words], dtype=np.float64)
    cdf = np.cumsum(freq / freq.sum())
    lines = []
    for _ in range(n_lines):
        k = int(rng.integers(min_words, max_words + 1))
        ids = np.minimum(np.searchsorted(cdf, rng.random(k)), len(words) - 1)
        lines.append(' '.join(words[int(i)][0] for i in ids))
    return lines


def write_test_corpus(root, lines):
    os.makedirs(os.path.join(root, 'data'), exist_ok=True)
    with open(os.path.join(root, 'data', 'test.txt'), 'w', encoding='utf-8') as f:
        f.write('\n'.join(lines) + '\n')


# --------------------------------------------------------------------------------------
# char-RNN experiments (reference decoder/decoder.py:244-341, train/data.py:28-46)
# --------------------------------------------------------------------------------------
def make_char_lexicon(vocab_size, seed=0, n_chars=48, extra_oov=200):
    """Lexicon whose display strings are 1-3 characters over a small alphabet, so that (a) most words
    take several character steps, (b) different segmentations of a reading spell the same string (the
    decoder de-duplicates paths by string, decoder.py:293-297) and (c) one reading carries the same
    display string under several parts of speech (per-substring de-duplication, decoder.py:113-122)."""
    rng = np.random.default_rng(seed)
    chars = [chr(0x4E00 + k) for k in range(n_chars)]
    n_pool = max(vocab_size // 3, 8)
    lens = rng.choice(np.arange(1, 7), size=n_pool, p=_LEN_P)
    pool, seen = [], set()
    for L in lens:
        r = ''.join(KANA[k] for k in rng.integers(0, len(KANA), size=int(L)))
        if r not in seen:
            seen.add(r)
            pool.append(r)
    n_words = vocab_size - 2 + extra_oov
    w = 1.0 / np.power(np.arange(1, len(pool) + 1), 0.6)
    w /= w.sum()
    ridx = rng.choice(len(pool), size=n_words, p=w)
    dlen = rng.choice([1, 2, 3], size=n_words, p=[.3, .45, .25])
    # characters are drawn from a vocabulary-dependent prefix of the alphabet for in-vocabulary words and from
    # the whole alphabet for the tail, so a few tail characters are unknown to the model (_char_check_oov)
    lexicon = [('<eos>', 10 ** 8)]
    reading_dict = {}
    used = set()
    kind = rng.random(size=n_words)
    pick = rng.integers(0, 1 << 30, size=(n_words, 2))
    base = []                                   # (display, reading) of the plain words generated so far
    for n in range(n_words):
        reading = pool[int(ridx[n])]
        hi = n_chars - 4 if n < vocab_size - 2 else n_chars
        disp = ''.join(chars[k] for k in rng.integers(0, hi, size=int(dlen[n])))
        if len(base) >= 8 and kind[n] < 0.2:
            # compound: spelled and read as two earlier words back to back -> the segmentations [a, b]
            # and [ab] give the same string
            (da, ra), (db, rb) = base[int(pick[n, 0]) % len(base)], base[int(pick[n, 1]) % len(base)]
            if len(da + db) <= 4 and len(ra + rb) <= 8:
                disp, reading = da + db, ra + rb
        elif len(base) >= 8 and kind[n] < 0.3:
            # homograph: an earlier word's spelling and reading under another part of speech
            disp, reading = base[int(pick[n, 0]) % len(base)]
        else:
            base.append((disp, reading))
        pos = 'P%d' % int(rng.integers(0, 3))
        word = '{}/{}/{}'.format(disp, reading, pos)
        if word in used:
            word = '{}/{}/Q{}'.format(disp, reading, n)
        used.add(word)
        lexicon.append((word, 10 ** 8 // (n + 1)))
        reading_dict.setdefault(reading, []).append(n + 1)
    return lexicon, reading_dict


def make_char_sentences(lexicon, n_sent, min_len=10, seed=1, vocab_size=None):
    """Kana strings from in-vocabulary words drawn UNIFORMLY (compounds and homographs are then common,
    so string de-duplication and the per-reading display filter are exercised in most frames)."""
    rng = np.random.default_rng(seed)
    words = lexicon[1:(vocab_size - 1 if vocab_size else len(lexicon))]
    out = []
    for _ in range(n_sent):
        s = ''
        while len(s) < min_len:
            s += words[int(rng.integers(0, len(words)))][0].split('/')[1]
        out.append(s)
    return out


def char_vocab(lexicon, vocab_size):
    """CharVocab (train/data.py:28-46): c2i over the display strings of lexicon[2:] of the size-limited
    vocabulary ('<unk>' 0, '<eos>' 1, then characters in order of first appearance)."""
    lex = [('<unk>', 0)] + list(lexicon[:vocab_size - 1])
    c2i = {'<unk>': 0, '<eos>': 1}
    for item in lex[2:]:
        for c in item[0].split('/')[0]:
            if c not in c2i:
                c2i[c] = len(c2i)
    return c2i


def make_char_experiment(root, experiment_id, vocab_size, hidden_size, embed_size, seed=0):
    """Tied standard-softmax character LM over CharVocab; config['vocab_size'] stays the WORD vocabulary
    size (it is what Vocab/CharVocab are built with), the model's output width is len(c2i)."""
    lexicon, reading_dict = make_char_lexicon(vocab_size, seed=seed)
    c2i = char_vocab(lexicon, vocab_size)
    cfg = make_config(len(c2i), hidden_size, embed_size, MODE_TIED)
    weights = make_weights(cfg, seed=seed)
    cfg['vocab_size'] = vocab_size
    cfg['char_rnn'] = True
    write_experiment(root, experiment_id, cfg, weights, lexicon, reading_dict)
    return cfg, weights, lexicon, reading_dict

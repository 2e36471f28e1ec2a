"""CPU-side checks (no GPU): host lattice / vocabulary logic against the reference-generated golden
fixtures, the packed CSR layout, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pytest

from jlm_b200 import _lib, lattice
from jlm_b200.vocab import Vocab
from oracle import jlm_oracle as O
from tests.golden.cases import CASES
from tests.helpers import build_case, load_golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('name', ['small_tied', 'small_tied_beam50', 'cfg2_tied'])
def test_lattice_matches_reference(name):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    meta, _ = load_golden(name)
    vocab = Vocab(cfg['vocab_size'], lexicon=lexicon)
    builder = lattice.LatticeBuilder(vocab.w2i, lexicon, reading_dict)
    for si, sent in enumerate(sentences):
        frames = builder.build(sent)
        g = meta['decode'][si]['lattice']
        assert len(frames) == len(sent) + 1
        for t, fr in enumerate(frames):
            assert [[n[0], n[1], n[2]] for n in fr] == g[str(t)], (name, si, t)
        # and against the oracle's independent restatement
        w2i, _ = O.make_vocab(lexicon, cfg['vocab_size'])
        assert frames == O.build_lattice(sent, w2i, lexicon, reading_dict)


def test_lattice_unk_fallback_and_oov_skip():
    # 3 in-vocab words + one OOV word (beyond vocab_size); kana 'ウ' has no reading at all
    lexicon = [('<eos>', 100), ('a/ア/P', 50), ('b/アイ/P', 40), ('c/イ/P', 30), ('oov/ア/P', 1)]
    reading = {'ア': [1, 4], 'アイ': [2], 'イ': [3]}
    vocab = Vocab(5, lexicon=lexicon)          # <unk>, <eos>, a, b, c  -> 'oov' is out of vocabulary
    frames = lattice.LatticeBuilder(vocab.w2i, lexicon, reading).build('アイウ')
    assert frames[0] == [(-1, 1, '<eos>')]
    assert frames[1] == [(0, 2, 'a/ア/P')]
    assert frames[2] == [(0, 3, 'b/アイ/P'), (1, 4, 'c/イ/P')]
    assert frames[3] == [(2, 0, 'ウ')]          # <unk> node carries the raw kana (quirk 8)
    bl = lattice.to_backward_lookup(frames)
    assert bl[2][0].start_idx == 0 and bl[2][0].reading_length == 2 and bl[0][0].reading_length == 1


@pytest.mark.parametrize('name', ['small_tied_vs', 'small_tied_vs_top', 'small_tied_vs_rand'])
def test_static_vocab_matches_reference(name):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    meta, _ = load_golden(name)
    vocab = Vocab(cfg['vocab_size'], lexicon=lexicon)
    builder = lattice.LatticeBuilder(vocab.w2i, lexicon, reading_dict)
    kw = case['decode_kwargs']
    for si, sent in enumerate(sentences):
        if kw.get('random_sampling'):
            np.random.seed(1234 + si)
        lv = lattice.static_vocab(builder.build(sent), len(vocab.w2i), kw.get('samples', 0),
                                  kw.get('top_sampling', False), kw.get('random_sampling', False))
        # the fixture stores the list as left after the LAST sentence only
        if si == len(sentences) - 1:
            assert lv == meta['decode'][si]['lattice_vocab']
        assert lv == sorted(set(lv))


@pytest.mark.parametrize('name', ['small_tied_dyn', 'small_tied_dyn_top', 'small_tied_dyn_rand'])
def test_dynamic_vocab_matches_reference(name):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    meta, _ = load_golden(name)
    vocab = Vocab(cfg['vocab_size'], lexicon=lexicon)
    builder = lattice.LatticeBuilder(vocab.w2i, lexicon, reading_dict)
    kw = case['decode_kwargs']
    for si, sent in enumerate(sentences):
        if kw.get('random_sampling'):
            np.random.seed(1234 + si)
        frames = builder.build(sent)
        lv0, news = lattice.dynamic_vocab(frames, len(vocab.w2i), kw.get('samples', 0),
                                          kw.get('top_sampling', False), kw.get('random_sampling', False))
        final = lattice.dynamic_vocab_final(lv0, news)
        assert {str(k): v for k, v in final.items()} == meta['decode'][si]['lattice_vocab']
        packed = lattice.PackedLattices([frames], dynamic=[(lv0, news)])
        # columns ordered by first appearance; lattice_vocab[i] == set(cols[:vfp[i+1]])
        cum = set(lv0)
        for i in range(len(frames)):
            if i:
                cum |= set(news[i])
            assert set(packed.vocab_ids[:packed.vocab_frame_ptr[i + 1]].tolist()) == cum
        ndup = len(lv0) - len(set(lv0))
        assert int(packed.dup_ptr[1]) == ndup


def test_packed_lattice_csr_layout():
    f0 = [[(-1, 1, '<eos>')], [(0, 5, 'x')], [(0, 6, 'y'), (1, 7, 'z')]]
    f1 = [[(-1, 1, '<eos>')], [(0, 9, 'q')]]
    p = lattice.PackedLattices([f0, f1])
    assert p.sent_len.tolist() == [2, 1]
    assert p.frame_ptr_off.tolist() == [0, 4]
    assert p.frame_ptr.tolist() == [0, 1, 2, 4, 4, 5, 6]
    assert p.node_start.tolist() == [-1, 0, 0, 1, -1, 0]
    assert p.node_word.tolist() == [1, 5, 6, 7, 1, 9]
    assert p.node_words(1) == ['<eos>', 'q']
    lb = p.c_struct()
    assert lb.n_sent == 2 and not lb.vocab_ptr


def test_library_exports_every_declared_symbol():
    """include/jlm_b200.h is the contract: every function it declares must be exported, with the
    binding table in jlm_b200/_lib.py covering exactly the same set."""
    hdr = open(os.path.join(REPO, 'include', 'jlm_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(jlm_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().jlm_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes implied by include/jlm_b200.h on LP64
    assert ctypes.sizeof(_lib.Config) == 4 * (6 + 3 * 8)
    assert ctypes.sizeof(_lib.Weights) == 8 * (4 + 4 + 4 + 4 + 8 + 8)
    assert ctypes.sizeof(_lib.LatticeBatch) == 8 * 11
    assert ctypes.sizeof(_lib.NBest) == 8 + 8 * 4
    assert ctypes.sizeof(_lib.BatchInfo) == 8 * 3 + 4 * 2 + 8 * 3 + 4 * 3 + 4 * 4 + 4 * 2 + 4 + 8 * 2 + 4 * 2


def test_product_fails_loudly_without_gpu(tmp_path):
    """No CPU fallback: constructing the model without a CUDA device must raise, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from jlm_b200 import synth, config, LSTM_Model
    synth.make_experiment(str(tmp_path), 3, 200, 64, 32, 'tied', seed=0)
    config.set_root(str(tmp_path))
    with pytest.raises(_lib.JlmError, match='no usable CUDA device'):
        LSTM_Model(3)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under jlm_b200/ may reference it."""
    pkg = os.path.join(REPO, 'jlm_b200')
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(root, fn), encoding='utf-8').read()
                assert 'oracle' not in src.replace('test infrastructure', ''), os.path.join(root, fn)


# ---------------------------------------------------------------------------------------------------
# native lattice builder (jlm_lattice_build, host C++; no GPU needed) vs the Python restatement of
# decoder.py:79-151 / decoder_dynamic.py:30-46 that is itself pinned to the reference fixtures above
# ---------------------------------------------------------------------------------------------------
def _native_setup(name):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    vocab = Vocab(cfg['vocab_size'], lexicon=lexicon)
    builder = lattice.LatticeBuilder(vocab.w2i, lexicon, reading_dict)
    nlex = lattice.NativeLexicon(vocab.w2i, lexicon, reading_dict)
    return case, vocab, builder, nlex, sentences


@pytest.mark.parametrize('name', ['small_tied', 'small_tied_beam50', 'cfg2_tied'])
def test_native_lattice_matches_python_builder(name):
    case, vocab, builder, nlex, sentences = _native_setup(name)
    texts = list(sentences) + [sentences[0][:1], sentences[-1][:3], 'ヰヱヰ']     # ragged + kana with no reading
    frames = [builder.build(t) for t in texts]
    want = lattice.PackedLattices(frames)
    got = lattice.NativeLattices(nlex, texts)
    for key in ('sent_len', 'frame_ptr_off', 'frame_ptr', 'node_start', 'node_word'):
        assert np.array_equal(getattr(got, key), getattr(want, key)), key
    for s in range(len(texts)):
        assert got.frames_of(s) == frames[s]
    lb = got.c_struct()
    assert lb.n_sent == len(texts) and not lb.vocab_ptr


@pytest.mark.parametrize('name,mode', [('small_tied_vs', 1), ('small_tied_vs_top', 1), ('small_tied_vs_rand', 1),
                                       ('small_tied_dyn', 2), ('small_tied_dyn_top', 2), ('small_tied_dyn_rand', 2),
                                       ('cfg4_tied_dyn', 2)])
def test_native_vocab_lists_match_python(name, mode):
    case, vocab, builder, nlex, sentences = _native_setup(name)
    kw = case['decode_kwargs']
    samples, top, rnd = kw.get('samples', 0), kw.get('top_sampling', False), kw.get('random_sampling', False)
    frames = [builder.build(t) for t in sentences]
    np.random.seed(99)
    if mode == 1:
        lists = [lattice.static_vocab(fr, len(vocab.w2i), samples, top, rnd) for fr in frames]
        want = lattice.PackedLattices(frames, vocab_lists=lists)
    else:
        dyn = [lattice.dynamic_vocab(fr, len(vocab.w2i), samples, top, rnd) for fr in frames]
        want = lattice.PackedLattices(frames, dynamic=dyn)
    # the same draws, in the same per-sentence order, for the native builder
    np.random.seed(99)
    extra = None
    if samples and rnd:
        extra = np.stack([np.random.randint(len(vocab.w2i), size=samples) for _ in sentences]).astype(np.int32)
    elif samples and top:
        extra = np.tile(np.arange(samples, dtype=np.int32), (len(sentences), 1))
    got = lattice.NativeLattices(nlex, sentences, mode, extra)
    v = got.c_struct()
    S = len(sentences)
    assert [int(v.vocab_ptr[i]) for i in range(S + 1)] == want.vocab_ptr.tolist()
    n = int(want.vocab_ptr[-1])
    assert [int(v.vocab_ids[i]) for i in range(n)] == want.vocab_ids.tolist()
    if mode == 2:
        nf = len(want.vocab_frame_ptr)
        assert [int(v.vocab_frame_ptr[i]) for i in range(nf)] == want.vocab_frame_ptr.tolist()
        assert [int(v.dup_ptr[i]) for i in range(S + 1)] == want.dup_ptr.tolist()
        nd = int(want.dup_ptr[-1])
        assert [int(v.dup_ids[i]) for i in range(nd)] == want.dup_ids.tolist()[:nd]
    else:
        assert got.vocab_list(S - 1) == lists[-1]


def test_native_lexicon_rejects_bad_input():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rptr = np.array([0, 1, 2], dtype=np.int64)
    chars = np.array([0x30A2, 0x30A2], dtype=np.uint32)          # the same reading twice
    wptr = np.array([0, 1, 2], dtype=np.int64)
    wids = np.array([2, 3], dtype=np.int32)
    rc = lib.jlm_lexicon_create(2, _lib.ptr(rptr, ctypes.c_int64), _lib.ptr(chars, ctypes.c_uint32),
                                _lib.ptr(wptr, ctypes.c_int64), _lib.ptr(wids, ctypes.c_int32), 1, 0, ctypes.byref(h))
    assert rc != 0 and b'duplicate reading' in lib.jlm_last_error()


def test_streaming_entry_points_reject_bad_arguments_without_a_device():
    """jlm_decode_texts_submit / _collect / _cancel validate before touching CUDA: null handles and jobs come back
    as status 1 with a message (never a crash), and cancel(NULL) is a no-op."""
    lib = _lib.load()
    job = ctypes.c_void_p()
    rc = lib.jlm_decode_texts_submit(None, None, 1, None, None, 5, 5, 0, 0, None, 0, 0, 0, ctypes.byref(job))
    assert rc != 0 and b'jlm_decode_texts_submit' in lib.jlm_last_error() and not job.value
    nb = _lib.TextNBest()
    rc = lib.jlm_decode_texts_collect(None, ctypes.byref(nb), None)
    assert rc != 0 and b'jlm_decode_texts_collect' in lib.jlm_last_error()
    assert lib.jlm_decode_texts_cancel(None) == 0
    assert lib.jlm_batch_fetch_async(None) != 0 and b'jlm_batch_fetch_async' in lib.jlm_last_error()
    rc = lib.jlm_decode_texts(None, None, 1, None, None, 5, 5, 0, 0, None, 0, 0, ctypes.byref(nb), None)
    assert rc != 0
    # state pool: same contract
    pool = ctypes.c_void_p()
    assert lib.jlm_pool_create(None, 16, ctypes.byref(pool)) != 0 and b'jlm_pool_create' in lib.jlm_last_error()
    assert lib.jlm_pool_step(None, 1, None, None, None) != 0 and lib.jlm_pool_nll(None, 1, None, None, None) != 0
    assert lib.jlm_pool_reset(None) != 0 and lib.jlm_pool_destroy(None) == 0


def test_public_header_is_plain_c(tmp_path):
    """include/jlm_b200.h is the FFI contract (cgo / JNI / ctypes bind against it): it must compile as C99 and as
    C++ on its own, with no CUDA or torch types."""
    import shutil
    import subprocess
    src = tmp_path / 'h.c'
    src.write_text('#include "jlm_b200.h"\nint main(void) { jlm_config c; (void)c; return (int)sizeof(jlm_nbest) == 0; }\n')
    inc = os.path.join(REPO, 'include')
    hdr = open(os.path.join(inc, 'jlm_b200.h')).read()
    assert re.findall(r'#include\s*[<"]([^>"]+)', hdr) == ['stdint.h']      # nothing but the C standard integer types
    if shutil.which('gcc'):
        subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-I', inc, str(src)], check=True)
    if shutil.which('g++'):
        subprocess.run(['g++', '-std=c++17', '-fsyntax-only', '-x', 'c++', '-I', inc, str(src)], check=True)


@pytest.mark.parametrize('mode', ['tied', 'untied', 'dsoftmax', 'dsoftmax_star'])
def test_text_weight_dumps_round_trip(mode, tmp_path):
    """train/weights.py:69-87 (--verbose): np.savetxt dumps per parameter -> the dict lstm_weights.pkl holds, bit for
    bit, for every projection mode (D-softmax keeps LM as a list of blocks dumped as LM0.txt, LM1.txt, ...)."""
    import numpy as np
    from jlm_b200 import synth, weights_io
    cfg = synth.make_config(300, 32, 16, mode)
    weights = synth.make_weights(cfg, seed=5)
    d = str(tmp_path / 'weights')
    weights_io.dump_text_weights(weights, d, with_npy=False)
    if mode == 'tied':
        assert open(os.path.join(d, 'embedding.txt')).readline().startswith('0 ')
    got = weights_io.load_text_weights(d, cfg)
    assert sorted(got) == sorted(weights)
    for k, v in weights.items():
        if isinstance(v, list):
            assert len(got[k]) == len(v)
            for a, b in zip(got[k], v):
                assert a.dtype == np.float32 and np.array_equal(a, b)
        else:
            assert got[k].dtype == np.float32 and got[k].shape == v.shape and np.array_equal(got[k], v), k
    w2 = weights_io.import_text_dump(d, cfg)
    import pickle
    back = pickle.load(open(os.path.join(d, 'lstm_weights.pkl'), 'rb'))
    assert sorted(back) == sorted(w2)

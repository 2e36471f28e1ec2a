"""Pins the CPU oracle (oracle/jlm_oracle.py) against outputs of the unmodified reference
(tests/golden/, produced by tests/golden/make_golden.py).  Everything here is float64 numpy on the
same shapes the reference uses, so agreement is expected to be bit-exact; a 1e-12 slack is allowed
only for the float sums that depend on BLAS kernel selection on a different host CPU."""
import numpy as np
import pytest

from oracle import jlm_oracle as O
from tests.golden.cases import CASES
from tests.helpers import build_case, load_golden, norm_paths

TOL = 1e-9
SMALL = [n for n in CASES if n.startswith('small')]
BIG = [n for n in CASES if not n.startswith('small')]


def _check_case(name):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    meta, arr = load_golden(name)
    dyn = case.get('dynamic', False)
    dec = O.OracleDecoder(cfg, weights, lexicon, reading_dict, dynamic=dyn)
    hs = meta['h_stride']
    for si, sent in enumerate(sentences):
        g = meta['decode'][si]
        if case['decode_kwargs'].get('random_sampling'):
            np.random.seed(1234 + si)
        trace, dyn_trace = [], []
        res = dec.decode(sent, trace=trace, dyn_trace=dyn_trace, **case['decode_kwargs'])
        # lattice: same nodes, same order (decoder.py:79-135)
        for t, fr in enumerate(dec.frames):
            assert [[n[0], n[1], n[2]] for n in fr] == g['lattice'][str(t)]
        # n-best: identical word sequences, scores to float64 round-off
        assert [ws for _, ws in res] == [ws for _, ws in g['nbest']]
        np.testing.assert_allclose([s for s, _ in res], [s for s, _ in g['nbest']], rtol=0, atol=TOL)
        # every LM step: same pruned paths in the same rank order
        assert len(trace) == len(g['steps'])
        for fi, (tr, gs) in enumerate(zip(trace, g['steps'])):
            a, b = norm_paths(tr['paths']), norm_paths(gs['paths'])
            assert [p[1] for p in a] == [p[1] for p in b], (name, si, fi)
            np.testing.assert_allclose([p[0] for p in a], [p[0] for p in b], rtol=0, atol=TOL)
            key = 's%d_f%d' % (si, fi)
            np.testing.assert_allclose(tr['lse'], arr[key + '_lse'], rtol=0, atol=TOL)
            np.testing.assert_allclose(tr['state'].sum(axis=1), arr[key + '_hsum'], rtol=0, atol=TOL)
            np.testing.assert_allclose(tr['cell'].sum(axis=1), arr[key + '_csum'], rtol=0, atol=TOL)
            if key + '_h' in arr:
                np.testing.assert_allclose(tr['state'][:, ::hs], arr[key + '_h'], rtol=0, atol=TOL)
                np.testing.assert_allclose(tr['cell'][:, ::hs], arr[key + '_c'], rtol=0, atol=TOL)
        if dyn:
            assert len(dyn_trace) == len(g['dyn_frames'])
            for fi, (tr, gs) in enumerate(zip(dyn_trace, g['dyn_frames'])):
                a, b = norm_paths(tr), norm_paths(gs)
                assert [p[1] for p in a] == [p[1] for p in b], (name, si, fi)
                np.testing.assert_allclose([p[0] for p in a], [p[0] for p in b], rtol=0, atol=TOL)
            assert {str(k): [int(x) for x in v] for k, v in dec.lattice_vocab.items()} == g['lattice_vocab']
        elif g['lattice_vocab'] is not None:
            assert [int(x) for x in dec.lattice_vocab] == g['lattice_vocab']


@pytest.mark.parametrize('name', SMALL)
def test_oracle_decode_matches_reference_small(name):
    _check_case(name)


@pytest.mark.parametrize('name', BIG)
def test_oracle_decode_matches_reference_full_size(name):
    _check_case(name)


@pytest.mark.parametrize('name', ['small_tied', 'small_untied', 'small_dsoftmax', 'small_dsoftmax_star',
                                  'small_tied_selfnorm', 'cfg2_tied', 'cfg3_dsoftmax_star', 'cfg5_dsoftmax_star',
                                  'small_tied_unsorted', 'small_dsoftmax_unsorted', 'small_dsoftmax_star_unsorted'])
def test_oracle_model_matches_reference(name):
    """LSTM_Model.predict_with_context / project (model.py:106-198) incl. vocab subsets."""
    case, cfg, weights, lexicon, reading_dict, _ = build_case(name)
    meta, arr = load_golden(name)
    m = O.OracleModel(cfg, weights)
    ys = meta['y_stride']
    probe = case['model_probe']
    B = len(probe['index'][0])
    h = np.zeros((B, m.hidden_size))
    c = np.zeros((B, m.hidden_size))
    for step, idx in enumerate(probe['index']):
        pred, y, h, c = m.predict(idx, h, c, None)
        np.testing.assert_allclose(pred[:, ::ys], arr['m_step%d_pred' % step], rtol=0, atol=TOL)
        np.testing.assert_allclose(y[:, ::ys], arr['m_step%d_y' % step], rtol=0, atol=TOL)
        np.testing.assert_allclose(h, arr['m_step%d_h' % step], rtol=0, atol=TOL)
        np.testing.assert_allclose(c, arr['m_step%d_c' % step], rtol=0, atol=TOL)
    if probe.get('vocab') is not None:
        if meta['model'] == ['vocab_ok']:
            yv = m.project(h, probe['vocab'])
            np.testing.assert_allclose(yv, arr['m_project_vocab_y'], rtol=0, atol=TOL)
            pv, yv2, _, _ = m.predict(probe['index'][-1], h, c, probe['vocab'])
            np.testing.assert_allclose(pv, arr['m_predict_vocab_pred'], rtol=0, atol=TOL)
            np.testing.assert_allclose(yv2, arr['m_predict_vocab_y'], rtol=0, atol=TOL)
        else:
            # quirk 2: untied + vocab raises in the reference (model.py:189)
            with pytest.raises(Exception):
                m.project(h, probe['vocab'])

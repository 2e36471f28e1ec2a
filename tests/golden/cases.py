"""Golden-fixture case table shared by make_golden.py (reference run) and the tests.

Small cases follow BASELINE.json configs[0] (V=1000 H=64 beam=5); the two large ones are
configs[1] and configs[2] (V=50000 H=512, standard tied / D-softmax*), two sentences each.
"""

_SMALL = dict(vocab_size=1000, hidden_size=64, embed_size=32, n_sent=4, min_len=12, seed=0,
              model_probe={'index': [[1, 5, 17], [2, 900, 33]], 'vocab': [1, 3, 7, 250, 400, 650, 999]})
_BEAM5 = dict(topN=10, beam_width=5)
_UNSORTED = [650, 3, 999, 1, 400, 7, 250, 2, 651]


def _c(base, **kw):
    d = dict(base)
    d.update(kw)
    return d


CASES = {
    # --- four projection modes, full softmax (decoder/decoder.py:220-241) ---
    'small_tied': _c(_SMALL, mode='tied', decode_kwargs=_BEAM5),
    'small_untied': _c(_SMALL, mode='untied', decode_kwargs=_BEAM5),
    'small_dsoftmax': _c(_SMALL, mode='dsoftmax', decode_kwargs=_BEAM5),
    'small_dsoftmax_star': _c(_SMALL, mode='dsoftmax_star', decode_kwargs=_BEAM5),
    # self-normalised model: pred = exp(y) (decoder/model.py:117-118)
    'small_tied_selfnorm': _c(_SMALL, mode='tied', self_norm=True, decode_kwargs=_BEAM5),
    # beam wider than most frames -> beams shorter than beam_width, topN cut
    'small_tied_beam50': _c(_SMALL, mode='tied', decode_kwargs=dict(topN=3, beam_width=50)),
    # static vocabulary selection (decoder/decoder.py:137-151)
    'small_tied_vs': _c(_SMALL, mode='tied', decode_kwargs=dict(_BEAM5, vocab_select=True)),
    'small_tied_vs_top': _c(_SMALL, mode='tied',
                            decode_kwargs=dict(_BEAM5, vocab_select=True, samples=20, top_sampling=True)),
    'small_tied_vs_rand': _c(_SMALL, mode='tied',
                             decode_kwargs=dict(_BEAM5, vocab_select=True, samples=20, random_sampling=True)),
    'small_dsoftmax_star_vs': _c(_SMALL, mode='dsoftmax_star', decode_kwargs=dict(_BEAM5, vocab_select=True)),
    'small_dsoftmax_vs': _c(_SMALL, mode='dsoftmax', decode_kwargs=dict(_BEAM5, vocab_select=True)),
    # incremental vocabulary selection (decoder/decoder_dynamic.py)
    'small_tied_dyn': _c(_SMALL, mode='tied', dynamic=True,
                         decode_kwargs=dict(_BEAM5, vocab_select=True)),
    'small_tied_dyn_top': _c(_SMALL, mode='tied', dynamic=True,
                             decode_kwargs=dict(_BEAM5, vocab_select=True, samples=20, top_sampling=True)),
    'small_tied_dyn_rand': _c(_SMALL, mode='tied', dynamic=True,
                              decode_kwargs=dict(_BEAM5, vocab_select=True, samples=20, random_sampling=True)),
    'small_tied_selfnorm_dyn': _c(_SMALL, mode='tied', self_norm=True, dynamic=True,
                                  decode_kwargs=dict(_BEAM5, vocab_select=True, samples=20, top_sampling=True)),
    # --- BASELINE.json configs[1] / configs[2] at full model size, 2 sentences ---
    'cfg2_tied': dict(vocab_size=50000, hidden_size=512, embed_size=256, mode='tied', n_sent=2, min_len=20,
                      seed=0, decode_kwargs=dict(topN=10, beam_width=10),
                      model_probe={'index': [[1, 7], [40000, 12]], 'vocab': [1, 2, 11999, 12000, 30000, 49999]}),
    'cfg3_dsoftmax_star': dict(vocab_size=50000, hidden_size=512, embed_size=256, mode='dsoftmax_star',
                               segments=[[200, 0, 12000], [100, 12000, 30000], [50, 30000, None]],
                               n_sent=2, min_len=20, seed=0, decode_kwargs=dict(topN=10, beam_width=10),
                               model_probe={'index': [[1, 7], [40000, 12]],
                                            'vocab': [1, 2, 11999, 12000, 30000, 49999]}),
    'cfg4_tied_dyn': dict(vocab_size=50000, hidden_size=512, embed_size=256, mode='tied', n_sent=2, min_len=20,
                          seed=0, dynamic=True,
                          decode_kwargs=dict(topN=10, beam_width=20, vocab_select=True, samples=200,
                                             top_sampling=True),
                          model_probe={'index': [[1, 7], [40000, 12]], 'vocab': None}),
    # --- BASELINE.json configs[4]: V=100000 H=1024 D-softmax*, the reference's default segments
    # (train/train.py:16,30), beam 50 (two kept-path registers per lane in k_prune_block, H=1024 gate tiles) ---
    'cfg5_dsoftmax_star': dict(vocab_size=100000, hidden_size=1024, embed_size=256, mode='dsoftmax_star',
                               segments=[[256, 0, 4000], [128, 4000, 12000], [64, 12000, None]],
                               n_sent=2, min_len=20, seed=0, decode_kwargs=dict(topN=10, beam_width=50),
                               model_probe={'index': [[1, 7], [90000, 12]],
                                            'vocab': [1, 2, 3999, 4000, 11999, 12000, 99999]}),
    # --- vocab subsets given UNSORTED (decoder/model.py:152-158,168-179): segmented projections
    # emit their columns segment-major while b2[vocab] keeps the caller's order; the tied softmax keeps list order ---
    'small_tied_unsorted': _c(_SMALL, mode='tied', n_sent=1, decode_kwargs=_BEAM5,
                              model_probe={'index': [[1, 5, 17], [2, 900, 33]], 'vocab': _UNSORTED}),
    'small_dsoftmax_unsorted': _c(_SMALL, mode='dsoftmax', n_sent=1, decode_kwargs=_BEAM5,
                                  model_probe={'index': [[1, 5, 17], [2, 900, 33]], 'vocab': _UNSORTED}),
    'small_dsoftmax_star_unsorted': _c(_SMALL, mode='dsoftmax_star', n_sent=1, decode_kwargs=_BEAM5,
                                       model_probe={'index': [[1, 5, 17], [2, 900, 33]], 'vocab': _UNSORTED}),
}

# --- char-RNN decoder (decoder/decoder.py:244-341): character LM over CharVocab, word lattice ---
_CHAR = dict(vocab_size=1000, hidden_size=64, embed_size=32, n_sent=4, min_len=10, seed=3)
CHAR_CASES = {
    'charrnn_small': _c(_CHAR, decode_kwargs=dict(topN=10, beam_width=5)),
    'charrnn_beam20': _c(_CHAR, n_sent=3, decode_kwargs=dict(topN=4, beam_width=20, vocab_select=True)),
}

#!/usr/bin/env python
"""Char-RNN decoder throughput (reference decoder/decoder.py:244-341): lock-step decode_batch over the device state
pool vs one decode() per sentence (predict_with_context round trips) vs the CPU oracle, same sentences.
Character LM: ~3000 characters, H=512, E=256; word lattice over a 20000-word lexicon; beam 10."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jlm_b200  # noqa: E402
from jlm_b200 import config, synth  # noqa: E402
from oracle import jlm_oracle as O  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
V, H, E, BEAM = 20000, 512, 256, 10
root = tempfile.mkdtemp(prefix='jlm_char_')
lexicon, reading_dict = synth.make_char_lexicon(V, seed=1, n_chars=3000)
c2i = synth.char_vocab(lexicon, V)
cfg = synth.make_config(len(c2i), H, E, synth.MODE_TIED)
weights = synth.make_weights(cfg, seed=1)
cfg['vocab_size'] = V
cfg['char_rnn'] = True
synth.write_experiment(root, 1, cfg, weights, lexicon, reading_dict)
config.set_root(root)
sents = synth.make_sentences(lexicon, S, min_len=20, seed=7, vocab_size=V)
chars = sum(len(s) for s in sents)
dec = jlm_b200.CharRNNDecoder(1)
dec.decode_batch(sents[:8], topN=BEAM, beam_width=BEAM)            # warm-up
t0 = time.perf_counter()
got = dec.decode_batch(sents, topN=BEAM, beam_width=BEAM)
t_batch = time.perf_counter() - t0
n1 = min(S, 24)
t0 = time.perf_counter()
one = [dec.decode(s, topN=BEAM, beam_width=BEAM) for s in sents[:n1]]
t_one = time.perf_counter() - t0
words, oc2i = O.make_char_vocab(lexicon, V)
model = O.OracleModel(cfg, weights)
n2 = min(S, 12)
t0 = time.perf_counter()
ora = []
for s in sents[:n2]:
    fr = O.build_lattice_char(s, words, oc2i, lexicon, reading_dict)
    ora.append(O.decode_charrnn(model, fr, oc2i, BEAM, BEAM))
t_ora = time.perf_counter() - t0
same = sum([w for _, w in a] == [w for _, w in b] for a, b in zip(got, ora))
c1 = sum(len(s) for s in sents[:n1])
c2 = sum(len(s) for s in sents[:n2])
print(json.dumps({'workload': 'char-RNN decoder: %d chars, H=%d, E=%d, beam %d, %d sentences (%d kana)' % (len(c2i), H, E, BEAM, S, chars),
                  'lockstep_pool_chars_per_s': chars / t_batch, 'per_sentence_chars_per_s': c1 / t_one,
                  'oracle_cpu_chars_per_s': c2 / t_ora, 'cpu_cores': os.cpu_count(),
                  'nbest_identical_to_oracle': '%d/%d' % (same, n2), 'pool_states': dec._pool.used}))

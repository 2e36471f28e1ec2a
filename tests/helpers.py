"""Shared test helpers: golden fixtures and regenerated synthetic experiments."""
import functools
import json
import os

import numpy as np

from jlm_b200 import synth
from tests.golden.cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@functools.lru_cache(maxsize=None)
def load_golden(name):
    meta = json.load(open(os.path.join(GOLDEN, name + '.json')))
    arrays = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    return meta, arrays


@functools.lru_cache(maxsize=4)
def build_case(name):
    """Regenerates the seeded experiment a fixture was produced from (checksum-verified)."""
    case = CASES[name]
    cfg = synth.make_config(case['vocab_size'], case['hidden_size'], case['embed_size'], case['mode'],
                            case.get('segments'), case.get('self_norm', False))
    weights = synth.make_weights(cfg, seed=case['seed'])
    lexicon, reading_dict = synth.make_lexicon(case['vocab_size'], seed=case['seed'])
    sentences = synth.make_sentences(lexicon, case['n_sent'], min_len=case['min_len'],
                                     seed=case['seed'] + 1, vocab_size=case['vocab_size'])
    meta, _ = load_golden(name)
    chk = float(sum(float(np.sum(np.asarray(v, dtype=np.float64))) for k, v in sorted(weights.items())
                    if not isinstance(v, list)))
    assert chk == meta['weights_checksum'], 'synthetic generator drifted from the golden fixture'
    assert sentences == meta['sentences']
    return case, cfg, weights, lexicon, reading_dict, sentences


def write_case(root, name, experiment_id=1):
    case, cfg, weights, lexicon, reading_dict, sentences = build_case(name)
    synth.write_experiment(root, experiment_id, cfg, weights, lexicon, reading_dict)
    return case, sentences


def norm_paths(paths):
    return [(float(s), [tuple(n) for n in nodes]) for s, nodes in paths]

#!/usr/bin/env python
"""Generates tests/golden/charrnn_*.json by running the reference's CharRNNDecoder (/root/reference,
decoder/decoder.py:244-341) on seeded synthetic character-LM experiments (jlm_b200/synth.py).

The class does not run as shipped: CharRNNDecoder._check_oov reads `self.vocab.words` and the lattice
indexes `self.w2i` with characters, but Decoder._load_vocab builds train.data.Vocab, which is word-keyed
and has no `.words` (AttributeError on the first reading that matches; the fixture records that outcome
under "as_shipped").  The fixtures are therefore produced by a subclass that overrides ONLY _load_vocab,
supplying the reference's own train.data.CharVocab (the vocabulary char models are trained with,
train/model.py:95-96) as `vocab`, its c2i/i2c as w2i/i2w, and `words`; every other method - lattice
construction, string de-duplication, multi-step word evaluation, pruning, the LM itself - is the
reference's unmodified code.  Dev container only (/root/reference does not exist on the GPU box).

    python tests/golden/make_golden_charrnn.py
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from tests.golden.cases import CHAR_CASES  # noqa: E402

WORKER = r'''
import sys, json, os, io, contextlib, traceback
sys.path.insert(0, '.')
import numpy as np
import warnings
warnings.simplefilter('ignore')
spec = json.load(open(sys.argv[1]))
with contextlib.redirect_stdout(io.StringIO()):
    import decoder as refdec
    from train.data import CharVocab

out = {'decode': []}
# 1. the class exactly as shipped
try:
    with contextlib.redirect_stdout(io.StringIO()):
        d0 = refdec.CharRNNDecoder(spec['experiment_id'])
        d0.decode(spec['sentences'][0], **spec['decode_kwargs'])
    out['as_shipped'] = 'ran'
except Exception as e:
    tb = traceback.extract_tb(sys.exc_info()[2])[-1]
    out['as_shipped'] = '%s: %s (decoder.py:%d)' % (type(e).__name__, e, tb.lineno)


# 2. with the vocabulary the class expects
class Shim(refdec.CharRNNDecoder):
    def _load_vocab(self):
        self.vocab = CharVocab(self.config['vocab_size'])
        self.vocab.words = set(w for w, _ in self.vocab.lexicon)
        self.i2w = self.vocab.i2c
        self.w2i = self.vocab.c2i


with contextlib.redirect_stdout(io.StringIO()):
    dec = Shim(spec['experiment_id'])
out['n_chars'] = len(dec.w2i)

rec = {'frames': [], 'calls': []}
orig_bp = dec._batch_predict
def bp(paths, vocab=None):
    rec['calls'].append(len(paths))
    orig_bp(paths, vocab)
dec._batch_predict = bp
orig_ef = dec._eval_frame
depth = [0]
def ef(paths):
    if depth[0] == 0:
        rec['frames'].append({'n_candidates': len(paths)})
    depth[0] += 1
    orig_ef(paths)
    depth[0] -= 1
dec._eval_frame = ef

for sent in spec['sentences']:
    rec['frames'], rec['calls'] = [], []
    # wrap decode's frame loop: record the pruned beam of every frame from the final _batch_predict call
    pruned = []
    def bp2(paths, vocab=None, _bp=bp):
        if depth[0] == 0:
            pruned.append([(float(p.neg_log_prob), [[int(n.start_idx), n.word] for n in p.nodes]) for p in paths])
        _bp(paths, vocab)
    dec._batch_predict = bp2
    res = dec.decode(sent, **spec['decode_kwargs'])
    dec._batch_predict = bp
    lat = {str(k): [[int(n.start_idx), int(n.word_idx), n.word] for n in v]
           for k, v in dec._last_lookup.items() if len(v)} if hasattr(dec, '_last_lookup') else None
    out['decode'].append({'input': sent, 'nbest': [[float(s), list(ws)] for s, ws in res],
                          'pruned': pruned, 'n_candidates': [f['n_candidates'] for f in rec['frames']],
                          'predict_rows': list(rec['calls'])})
# the lattice (CharRNNDecoder.decode keeps it local): rebuild through the same method
for d in out['decode']:
    bl = dec._build_lattice(d['input'])
    d['lattice'] = {str(k): [[int(n.start_idx), int(n.word_idx), n.word] for n in v] for k, v in bl.items() if len(v)}
json.dump(out, open(sys.argv[2], 'w'), ensure_ascii=False)
'''


def main():
    import numpy as np
    from jlm_b200 import synth
    scratch = tempfile.mkdtemp(prefix='jlm_ref_char_')
    ref = os.path.join(scratch, 'ref')
    shutil.copytree('/root/reference', ref)
    worker = os.path.join(scratch, 'worker.py')
    with open(worker, 'w') as f:
        f.write(WORKER)
    for name, case in CHAR_CASES.items():
        print('golden case', name, flush=True)
        for sub in ('data', os.path.join('train', 'experiments')):
            shutil.rmtree(os.path.join(ref, sub), ignore_errors=True)
        cfg, weights, lexicon, reading_dict = synth.make_char_experiment(
            ref, 1, case['vocab_size'], case['hidden_size'], case['embed_size'], seed=case['seed'])
        sentences = synth.make_char_sentences(lexicon, case['n_sent'], min_len=case['min_len'], seed=case['seed'] + 1,
                                              vocab_size=case['vocab_size'])
        spec = {'experiment_id': 1, 'sentences': sentences, 'decode_kwargs': case['decode_kwargs']}
        spec_path = os.path.join(scratch, 'spec.json')
        json.dump(spec, open(spec_path, 'w'), ensure_ascii=False)
        out_json = os.path.join(HERE, name + '.json')
        subprocess.run([sys.executable, worker, spec_path, out_json], cwd=os.path.join(ref, 'decoder'), check=True,
                       env=dict(os.environ, PYTHONWARNINGS='ignore'))
        meta = json.load(open(out_json))
        meta['case'] = case
        meta['sentences'] = sentences
        meta['weights_checksum'] = float(sum(float(np.sum(np.asarray(v, dtype=np.float64)))
                                             for k, v in sorted(weights.items()) if not isinstance(v, list)))
        json.dump(meta, open(out_json, 'w'), ensure_ascii=False)
        print('  as shipped:', meta['as_shipped'])
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main()

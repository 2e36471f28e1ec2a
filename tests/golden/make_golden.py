#!/usr/bin/env python
"""Generates tests/golden/*.json(+.npz) by running the UNMODIFIED reference (/root/reference).

The reference has no tests and ships no data, so parity is pinned on outputs
of the reference itself, run in this container on seeded synthetic experiments
(jlm_b200/synth.py). The reference derives every path from the location of its own config.py
(config.py:14-19) and /root/reference is read-only, so it is copied to a scratch directory
first; nothing from it is copied into this repository - only its OUTPUTS are committed.

Run (dev container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py            # regenerates every fixture
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

from tests.golden.cases import CASES  # noqa: E402

WORKER = r'''
import sys, json, os
sys.path.insert(0, '.')
import numpy as np
import warnings
warnings.simplefilter('ignore')
import io, contextlib
spec = json.load(open(sys.argv[1]))
with contextlib.redirect_stdout(io.StringIO()):
    import decoder as refdec
    import decoder_dynamic as refdyn
    from model import LSTM_Model

def path_nodes(p):
    return [[int(n.start_idx), int(n.word_idx)] for n in p.nodes]

out = {'decode': [], 'model': None}
BIG = spec['vocab_size'] > 5000
HS = 4 if BIG else 1          # column stride of stored h/c
YS = 97 if BIG else 1         # column stride of stored full-vocab pred/y
out['h_stride'], out['y_stride'] = HS, YS
arrays = {}
with contextlib.redirect_stdout(io.StringIO()):
    dec = (refdyn.DynamicDecoder if spec['dynamic'] else refdec.Decoder)(spec['experiment_id'])

rec = {'frames': []}
orig_bp = dec._batch_predict
def bp(paths, vocab=None):
    pre = [(float(p.neg_log_prob), path_nodes(p)) for p in paths]
    orig_bp(paths, vocab)
    ent = {'paths': pre, 'n_vocab': (len(vocab) if vocab is not None else None)}
    ent['state'] = np.concatenate([p.state for p in paths], axis=0)
    ent['cell'] = np.concatenate([p.cell for p in paths], axis=0)
    lg = np.stack([np.asarray(p.logits, dtype=np.float64) for p in paths], axis=0)
    m = lg.max(axis=1, keepdims=True)
    ent['lse'] = (m[:, 0] + np.log(np.exp(lg - m).sum(axis=1)))
    rec['frames'].append(ent)
dec._batch_predict = bp

rec_dyn = {'frames': []}
if spec['dynamic']:
    orig_bcf = dec._build_current_frame
    def bcf(frame, i, beam_width):
        orig_bcf(frame, i, beam_width)
        rec_dyn['frames'].append([(float(p.neg_log_prob), path_nodes(p)) for p in frame[i]])
    dec._build_current_frame = bcf

for si, sent in enumerate(spec['sentences']):
    rec['frames'] = []
    rec_dyn['frames'] = []
    if spec['decode_kwargs'].get('random_sampling'):
        np.random.seed(spec['np_seed'] + si)
    res = dec.decode(sent, **spec['decode_kwargs'])
    d = {'input': sent,
         'nbest': [[float(s), list(ws)] for s, ws in res],
         'lattice': {str(k): [[int(n.start_idx), int(n.word_idx), n.word] for n in v]
                     for k, v in dec.backward_lookup.items() if len(v)},
         'lattice_vocab': None, 'steps': [], 'dyn_frames': rec_dyn['frames']}
    lv = dec.lattice_vocab
    if isinstance(lv, list):
        d['lattice_vocab'] = [int(x) for x in lv]
    elif isinstance(lv, dict):
        d['lattice_vocab'] = {str(k): [int(x) for x in v] for k, v in lv.items()}
    for fi, ent in enumerate(rec['frames']):
        d['steps'].append({'paths': ent['paths'], 'n_vocab': ent['n_vocab']})
        key = 's%d_f%d' % (si, fi)
        arrays[key + '_lse'] = ent['lse']
        arrays[key + '_hsum'] = ent['state'].sum(axis=1)
        arrays[key + '_csum'] = ent['cell'].sum(axis=1)
        if fi < 3 or fi == len(rec['frames']) - 1:
            arrays[key + '_h'] = ent['state'][:, ::HS]
            arrays[key + '_c'] = ent['cell'][:, ::HS]
    out['decode'].append(d)

# model-level API: predict / project (decoder/model.py:106-123,141-193)
m = dec.model
mspec = spec['model_probe']
mo = []
idx0 = mspec['index'][0]
h = np.zeros((len(idx0), m.hidden_size)); c = np.zeros((len(idx0), m.hidden_size))
for step, idx in enumerate(mspec['index']):
    (pred, y, _, _), h, c = m.predict_with_context(idx, h, c, None)
    arrays['m_step%d_pred' % step] = pred[:, ::YS]
    arrays['m_step%d_y' % step] = y[:, ::YS]
    arrays['m_step%d_lse' % step] = y.max(axis=1) + np.log(np.exp(y - y.max(axis=1, keepdims=True)).sum(axis=1))
    arrays['m_step%d_h' % step] = h
    arrays['m_step%d_c' % step] = c
if mspec.get('vocab') is not None:
    try:
        yv = m.project(h, mspec['vocab'])
        arrays['m_project_vocab_y'] = yv
        (predv, yv2, _, _), h2, c2 = m.predict_with_context(mspec['index'][-1], h, c, mspec['vocab'])
        arrays['m_predict_vocab_pred'] = predv
        arrays['m_predict_vocab_y'] = yv2
        mo.append('vocab_ok')
    except Exception as e:
        mo.append('vocab_error:' + type(e).__name__)
out['model'] = mo
json.dump(out, open(sys.argv[2], 'w'), ensure_ascii=False)
np.savez(sys.argv[3], **arrays)
'''


def main():
    from jlm_b200 import synth
    only = set(sys.argv[1:])
    scratch = tempfile.mkdtemp(prefix='jlm_ref_')
    ref = os.path.join(scratch, 'ref')
    shutil.copytree('/root/reference', ref)
    os.makedirs(os.path.join(ref, 'decoder', 'eval'), exist_ok=True)
    worker = os.path.join(scratch, 'worker.py')
    with open(worker, 'w') as f:
        f.write(WORKER)
    for name, case in CASES.items():
        if only and name not in only:
            continue
        print('golden case', name, flush=True)
        for sub in ('data', os.path.join('train', 'experiments')):
            shutil.rmtree(os.path.join(ref, sub), ignore_errors=True)
        cfg, weights, lexicon, reading_dict = synth.make_experiment(
            ref, 1, case['vocab_size'], case['hidden_size'], case['embed_size'], case['mode'],
            segments=case.get('segments'), self_norm=case.get('self_norm', False), seed=case['seed'])
        sentences = synth.make_sentences(lexicon, case['n_sent'], min_len=case['min_len'],
                                         seed=case['seed'] + 1, vocab_size=case['vocab_size'])
        spec = {'experiment_id': 1, 'dynamic': case.get('dynamic', False), 'sentences': sentences,
                'decode_kwargs': case['decode_kwargs'], 'np_seed': 1234, 'vocab_size': case['vocab_size'],
                'model_probe': case['model_probe']}
        spec_path = os.path.join(scratch, 'spec.json')
        json.dump(spec, open(spec_path, 'w'), ensure_ascii=False)
        out_json = os.path.join(HERE, name + '.json')
        out_npz = os.path.join(HERE, name + '.npz')
        env = dict(os.environ, PYTHONWARNINGS='ignore')
        subprocess.run([sys.executable, worker, spec_path, out_json, out_npz],
                       cwd=os.path.join(ref, 'decoder'), check=True, env=env)
        # record what produced the fixture so the tests can regenerate identical inputs
        meta = json.load(open(out_json))
        meta['case'] = case
        meta['sentences'] = sentences
        import numpy as np
        meta['weights_checksum'] = float(sum(float(np.sum(np.asarray(v, dtype=np.float64)))
                                             for k, v in sorted(weights.items())
                                             if not isinstance(v, list)))
        json.dump(meta, open(out_json, 'w'), ensure_ascii=False)
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main()

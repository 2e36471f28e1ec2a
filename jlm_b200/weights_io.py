"""Weight import from the text dumps train/weights.py writes with --verbose (train/weights.py:66-87).

`dump_weights(weights_dict, dump_dir, verbose=True)` leaves, beside `lstm_weights.pkl`, one `np.savetxt` file per
parameter (`HMi.txt`, ..., `b2.txt`, `PM.txt` / `UM.txt`, `VT{i}.txt`, `LM{i}.txt`; the D-softmax list `LM` as
`LM0.txt`, `LM1.txt`, ...), a `.npy` twin for the plain arrays and `embedding.txt` (LM rows prefixed by the word id).
An experiment that only kept those dumps can be replayed: `load_text_weights` rebuilds the dict LSTM_Model reads from
the pickle (decoder/model.py:73-104), bit for bit - savetxt's '%.18e' round-trips float32.
"""
import os
import pickle

import numpy as np

GATE_NAMES = ['HMi', 'HMf', 'HMo', 'HMg', 'IMi', 'IMf', 'IMo', 'IMg', 'bi', 'bf', 'bo', 'bg', 'b2']


def parameter_names(config):
    """The variable list of train/weights.py:30-43 for this config (its duplicated 'HMo' entry dropped)."""
    names = list(GATE_NAMES)
    names.append('PM' if config.get('share_embedding', True) else 'UM')
    if config.get('V_table'):
        for i in range(len(config['embedding_seg'])):
            if i != 0:
                names.append('VT{}'.format(i))
            names.append('LM{}'.format(i))
    else:
        names.append('LM')
    return names


def dump_text_weights(weights, dump_dir, with_npy=True):
    """What dump_weights(..., verbose=True) writes (train/weights.py:69-87), for round-trip tests and for handing
    an experiment to tools that read the text form."""
    os.makedirs(dump_dir, exist_ok=True)
    for name, m in weights.items():
        if isinstance(m, list):
            for i, item in enumerate(m):
                np.savetxt(os.path.join(dump_dir, '{}{}.txt'.format(name, i)), item)
        else:
            np.savetxt(os.path.join(dump_dir, name + '.txt'), m)
            if with_npy:
                np.save(os.path.join(dump_dir, name + '.npy'), m)
    lm = os.path.join(dump_dir, 'LM.txt')
    if os.path.exists(lm):      # build_embedding_with_word (train/weights.py:89-97)
        with open(os.path.join(dump_dir, 'embedding.txt'), 'w') as out, open(lm) as f:
            for i, line in enumerate(f):
                out.write('{} {}'.format(i, line))


def _read(dump_dir, name):
    npy = os.path.join(dump_dir, name + '.npy')
    if os.path.exists(npy):
        return np.load(npy)
    txt = os.path.join(dump_dir, name + '.txt')
    if not os.path.exists(txt):
        raise FileNotFoundError('weight dump {} has neither {}.txt nor {}.npy'.format(dump_dir, name, name))
    a = np.loadtxt(txt, dtype=np.float64, ndmin=1)
    return a.astype(np.float32)


def load_text_weights(dump_dir, config):
    """The weights dict of lstm_weights.pkl rebuilt from the per-parameter text (or .npy) dumps."""
    w = {}
    for name in parameter_names(config):
        if name == 'LM' and config.get('D_softmax') and not os.path.exists(os.path.join(dump_dir, 'LM.txt')) \
                and not os.path.exists(os.path.join(dump_dir, 'LM.npy')):
            # D-softmax keeps LM as a list of per-segment blocks, dumped as LM0.txt, LM1.txt, ... (weights.py:46-55,76-80)
            w['LM'] = [_read(dump_dir, 'LM{}'.format(i)) for i in range(len(config['embedding_seg']))]
            continue
        a = _read(dump_dir, name)
        if name in ('bi', 'bf', 'bo', 'bg', 'b2'):
            a = a.reshape(-1)
        elif a.ndim == 1:
            a = a.reshape(1, -1) if name.startswith('IM') or name.startswith('HM') else a.reshape(-1, 1)
        w[name] = a
    return w


def import_text_dump(dump_dir, config, pickle_path=None):
    """Text dump -> lstm_weights.pkl (so that LSTM_Model, and the reference's own loader, can read the experiment)."""
    w = load_text_weights(dump_dir, config)
    if pickle_path is None:
        pickle_path = os.path.join(dump_dir, 'lstm_weights.pkl')
    with open(pickle_path, 'wb') as f:
        pickle.dump(w, f)
    return w

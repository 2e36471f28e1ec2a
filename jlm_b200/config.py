"""Directory layout shared with the reference (config.py:14-19).

The reference derives every path from the location of its own config.py.  This package keeps the
same layout under a configurable root:

    <root>/data/{lexicon.pkl, reading_dict.pkl}
    <root>/train/experiments/<id>/{config.json, weights/lstm_weights[_comp_N].pkl}

The root is $JLM_ROOT (default: current directory) and can be changed with set_root().
"""
import json
import os

root_path = os.path.abspath(os.environ.get('JLM_ROOT', os.getcwd()))
train_path = os.path.join(root_path, 'train')
data_path = os.path.join(root_path, 'data')
experiment_path = os.path.join(train_path, 'experiments')


def set_root(path):
    """Point the package at another experiment tree (same role as moving the reference's config.py)."""
    global root_path, train_path, data_path, experiment_path
    root_path = os.path.abspath(path)
    train_path = os.path.join(root_path, 'train')
    data_path = os.path.join(root_path, 'data')
    experiment_path = os.path.join(train_path, 'experiments')


class ExperimentConfig:
    """config.py:21-26"""

    def __init__(self, **entries):
        self.__dict__.update(entries)

    def __repr__(self):
        return str(self.__dict__)


def load_config(experiment_id):
    with open(os.path.join(experiment_path, str(experiment_id), 'config.json'), 'rt') as f:
        return json.loads(f.read())


def get_configs(experiment):
    """config.py:28-30"""
    return ExperimentConfig(**load_config(experiment))

"""Vocabulary (reference train/data.py:15-26): '<unk>' is id 0, then lexicon[:size-1]."""
import os
import pickle

from . import config


class Vocab(object):
    def __init__(self, size, lexicon=None):
        if lexicon is None:
            with open(os.path.join(config.data_path, 'lexicon.pkl'), 'rb') as f:
                lexicon = pickle.load(f)
        self.lexicon = [('<unk>', 0)] + list(lexicon[:size - 1])
        self.w2i = {x[0]: i for i, x in enumerate(self.lexicon)}
        self.i2w = {v: k for k, v in self.w2i.items()}

    def __len__(self):
        return len(self.w2i)

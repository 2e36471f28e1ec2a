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


class CharVocab(Vocab):
    """Character index over the display strings of the size-limited vocabulary (reference
    train/data.py:28-46): '<unk>' 0, '<eos>' 1, then characters in order of first appearance in lexicon[2:].
    `words` is the set CharRNNDecoder._check_oov tests membership in (decoder/decoder.py:263-264)."""

    def __init__(self, size, lexicon=None):
        super(CharVocab, self).__init__(size, lexicon)
        self.c2i = {'<unk>': 0, '<eos>': 1}
        for item in self.lexicon[2:]:
            for c in item[0].split('/')[0]:
                if c not in self.c2i:
                    self.c2i[c] = len(self.c2i)
        self.i2c = {v: k for k, v in self.c2i.items()}
        self.words = set(x[0] for x in self.lexicon)

    def __len__(self):
        return len(self.c2i)

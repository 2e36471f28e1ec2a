"""jlm_b200: B200-native (sm_100a) implementation of JLM's numpy inference hot path.

    from jlm_b200 import config, LSTM_Model, Decoder, DynamicDecoder
    config.set_root('/path/with/data/and/train/experiments')
    Decoder(experiment_id=27).decode('キョーワイーテンキデス', beam_width=10)

Mirrors reference decoder/model.py (LSTM_Model), decoder/decoder.py (Decoder) and
decoder/decoder_dynamic.py (DynamicDecoder) and decoder/decoder.py:244-341 (CharRNNDecoder).  The arithmetic runs in libjlm_b200.so (hand-written
CUDA, C ABI in include/jlm_b200.h); there is no CPU fallback.
"""
from . import config  # noqa: F401
from .model import LSTM_Model  # noqa: F401
from .decoder import Decoder  # noqa: F401
from .decoder_dynamic import DynamicDecoder  # noqa: F401
from .decoder_charrnn import CharRNNDecoder  # noqa: F401
from .vocab import Vocab, CharVocab  # noqa: F401

__all__ = ['config', 'LSTM_Model', 'Decoder', 'DynamicDecoder', 'CharRNNDecoder', 'Vocab', 'CharVocab']

"""Evaluation harness: drop-in for the reference's decoder/eval.py (accuracy + timing log).

Same command line (decoder/eval.py:17-30, including its `type=bool` flags for which any non-empty
string means True), same evaluation set selection (`load_eval_set`, eval.py:125-166), same hit
counting and the same log file name and contents (eval.py:50-122).  Differences: the n-gram and
char-RNN decoders are out of scope, and with `--batch` (default) all pairs are
decoded in one lock-step `decode_batch` call instead of one `decode` call per pair.

    python -m jlm_b200.eval -e 1 -es 100 -b 10 [--root DIR] [--device 0] [--no-batch]
"""
import argparse
import os
import sys
import time

import numpy as np

from . import config
from .decoder import Decoder
from .decoder_charrnn import CharRNNDecoder
from .decoder_dynamic import DynamicDecoder
from .vocab import Vocab


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument("--experiment_id", "-e", type=int, default=1, help="experiment id to eval")
    parser.add_argument("--eval_size", "-es", type=int, default=100, help="Number of sentences to evaluate")
    parser.add_argument("--use_ngram", "-ng", type=bool, default=False, help="Use ngram decoder or not")
    parser.add_argument("--ngram_order", "-o", type=int, default=3, help="Ngram order")
    parser.add_argument("--comp", "-c", type=int, default=0, help="Compression bit, 0 means no compression")
    parser.add_argument("--vocab_select", "-vs", type=bool, default=False, help="Use vocab select method or not")
    parser.add_argument("--top_sampling", "-ts", type=bool, default=False, help="Sampling strategy for vocab select")
    parser.add_argument("--random_sampling", "-rs", type=bool, default=False, help="Sampling strategy for vocab select")
    parser.add_argument("--samples", "-s", type=int, default=0, help="Samples when using advanced sampling")
    parser.add_argument("--beam_size", "-b", type=int, default=10, help="Beam size for decoder")
    parser.add_argument("--dynamic_decoding", "-dd", type=bool, default=False, help="Use incremental decoding or not")
    # additions of this implementation
    parser.add_argument("--root", default=None, help="directory holding data/ and train/ (default: $JLM_ROOT or cwd)")
    parser.add_argument("--device", type=int, default=0, help="CUDA device")
    parser.add_argument("--no-batch", dest="batch", action="store_false",
                        help="one decode() call per pair, as the reference does, instead of one lock-step batch")
    parser.add_argument("--log_dir", default="eval", help="directory of the eval_log_*.txt file")
    return parser


def log_name(args, decoder_type="neural"):
    """eval.py:65-76"""
    return 'eval_log_{}_e_{}_dynamic_{}_size_{}_b_{}_comp_{}_vocab_sel_{}_samples_{}_top_{}_random_{}.txt'.format(
        decoder_type, args.experiment_id, args.dynamic_decoding, args.eval_size, args.beam_size, args.comp,
        args.vocab_select, args.samples, args.top_sampling, args.random_sampling)


class Evaluator(object):
    def __init__(self, args, decoder=None):
        # eval.py:33-48
        self.args = args
        if args.use_ngram:
            raise NotImplementedError('the n-gram baseline decoder is out of scope (SURVEY.md section 2, row 6)')
        if decoder is not None:
            self.decoder = decoder
            self.config = getattr(decoder, 'config', {})
        else:
            self.config = config.load_config(args.experiment_id)
            if self.config.get('char_rnn'):
                cls = CharRNNDecoder                      # eval.py:43-44
            else:
                cls = DynamicDecoder if args.dynamic_decoding else Decoder
            self.decoder = cls(experiment_id=args.experiment_id, comp=args.comp, device=args.device)
        self.vocab = getattr(self.decoder, 'vocab', None) or Vocab(self.config['vocab_size'])
        self.w2i = self.vocab.w2i

    def load_eval_set(self):
        """eval.py:125-166: (reading, target) pairs from data/test.txt, lines with an OOV token dropped."""
        args = self.args
        x, y = [], []
        with open(os.path.join(config.data_path, 'test.txt'), 'r', encoding='utf-8') as f:
            lines = f.readlines()
        print('take {} for evaluation from all {} lines'.format(args.eval_size, len(lines)))
        for line in lines:
            tokens = line.strip().split(' ')
            if any(self.decoder._check_oov(t) for t in tokens):
                continue
            readings = ''.join([t.split('/')[1] if t.split('/')[1] != '' else t.split('/')[0] for t in tokens])
            target = ''.join([t.split('/')[0] for t in tokens])
            x.append(readings)
            y.append(target)
            if len(x) >= args.eval_size:
                break
        print('{} pairs load'.format(len(x)))
        return x, y

    def _decode_all(self, x_):
        args = self.args
        kw = dict(beam_width=args.beam_size, vocab_select=args.vocab_select, samples=args.samples,
                  top_sampling=args.top_sampling, random_sampling=args.random_sampling)
        if args.batch and hasattr(self.decoder, 'decode_batch') and x_:
            return self.decoder.decode_batch(x_, **kw)
        return [self.decoder.decode(x, **kw) for x in x_]

    def evaluate(self):
        """eval.py:50-122; returns (best_hit, n_best_hit, no_hit, pairs evaluated)."""
        args = self.args
        best_hit = n_best_hit = 0
        os.makedirs(args.log_dir, exist_ok=True)
        path = os.path.join(args.log_dir, log_name(args))
        with open(path, 'w', encoding='utf-8') as f:
            x_, y_ = self.load_eval_set()
            start_time = time.time()
            all_results = self._decode_all(x_)
            for x, y, results in zip(x_, y_, all_results):
                sentences = [''.join([w.split('/')[0] for w in item[1]]) for item in results]
                if y == sentences[0]:
                    best_hit += 1
                    f.write('best hit\n')
                elif y in sentences:
                    f.write('nbest hit\n')
                    n_best_hit += 1
                else:
                    f.write('no hit\n')
                f.write('{}\t{}\n'.format(y, x))
                for item in sentences:
                    f.write('{}\n'.format(item))
            summary = 'best_hit {} nbest_hit{} no_hit {} eval_size {}'.format(
                best_hit, n_best_hit, args.eval_size - best_hit - n_best_hit, args.eval_size)
            f.write(summary)
            d = self.decoder
            lstm = np.mean(d.perf_log_lstm) if len(d.perf_log_lstm) else float('nan')
            soft = np.mean(d.perf_log_softmax) if len(d.perf_log_softmax) else float('nan')
            per_sent = (np.sum(list(d.perf_log_lstm) + list(d.perf_log_softmax)) / d.perf_sen) if d.perf_sen else float('nan')
            f.write("--- %f seconds lstm per step ---" % lstm)
            f.write("--- %f seconds softmax per step ---" % soft)
            f.write("--- %f seconds per sent.---" % per_sent)
            f.write("--- %s seconds ---" % (time.time() - start_time))
            print(summary)
            print("--- %f seconds lstm per step ---" % lstm)
            print("--- %f seconds softmax per step ---" % soft)
            print("--- %f seconds per sent.---" % per_sent)
            if args.dynamic_decoding and getattr(d, 'perf_log_fix_vocab', None):
                print("--- %f seconds per step for vocab fix.---" % np.mean(d.perf_log_fix_vocab))
                print("--- %f seconds per step for lattice path fix.---" % np.mean(d.perf_log_fix_lattice_path_prob))
            print("--- %f seconds ---" % (time.time() - start_time))
        self.log_path = path
        return best_hit, n_best_hit, len(x_) - best_hit - n_best_hit, len(x_)


def parse_log(root='./'):
    """eval.py:168-178"""
    for folder, _, files in os.walk(root):
        for filename in files:
            if 'eval_log' in filename:
                print(filename)
                with open(os.path.join(folder, filename), 'r', encoding='utf-8') as f:
                    for line in f.readlines():
                        if 'best_hit' in line:
                            print(line.strip())


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.root:
        config.set_root(args.root)
    Evaluator(args).evaluate()


if __name__ == '__main__':
    main(sys.argv[1:])

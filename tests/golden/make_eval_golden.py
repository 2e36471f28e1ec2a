#!/usr/bin/env python
"""Generates tests/golden/eval_*.txt by running the UNMODIFIED reference decoder/eval.py on a seeded
synthetic experiment + data/test.txt (dev container only; nothing of the reference is copied into the
repository, only the log it writes).  The timing tail of the log is cut off.

    python tests/golden/make_eval_golden.py
"""
import glob
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

# name -> eval.py command line (decoder/eval.py:17-30); every flag is type=bool there, so '1' means True
EVAL_CASES = {
    'eval_static_b5': ['-e', '1', '-es', '12', '-b', '5'],
    'eval_dynamic_b5_top20': ['-e', '1', '-es', '12', '-b', '5', '-dd', '1', '-vs', '1', '-s', '20', '-ts', '1'],
    'eval_vocab_select_b3': ['-e', '1', '-es', '8', '-b', '3', '-vs', '1'],
}
EXPERIMENT = dict(vocab_size=1000, hidden_size=64, embed_size=32, mode='tied', seed=0, corpus_lines=60, corpus_seed=2)


def strip_timing(text):
    """Keeps everything up to and including 'eval_size N' (the reference appends the timing lines
    without a newline, eval.py:100-109)."""
    m = re.search(r'best_hit \d+ nbest_hit\d+ no_hit \d+ eval_size \d+', text)
    return text[:m.end()] + '\n'


def main():
    from jlm_b200 import synth
    scratch = tempfile.mkdtemp(prefix='jlm_ref_eval_')
    ref = os.path.join(scratch, 'ref')
    shutil.copytree('/root/reference', ref)
    os.makedirs(os.path.join(ref, 'decoder', 'eval'), exist_ok=True)
    e = EXPERIMENT
    cfg, weights, lexicon, reading = synth.make_experiment(ref, 1, e['vocab_size'], e['hidden_size'], e['embed_size'],
                                                           e['mode'], seed=e['seed'])
    synth.write_test_corpus(ref, synth.make_test_corpus(lexicon, e['corpus_lines'], seed=e['corpus_seed']))
    for name, argv in EVAL_CASES.items():
        for f in glob.glob(os.path.join(ref, 'decoder', 'eval', '*.txt')):
            os.remove(f)
        subprocess.run([sys.executable, 'eval.py'] + argv, cwd=os.path.join(ref, 'decoder'), check=True,
                       env=dict(os.environ, PYTHONWARNINGS='ignore'), stdout=subprocess.DEVNULL)
        logs = glob.glob(os.path.join(ref, 'decoder', 'eval', 'eval_log_*.txt'))
        assert len(logs) == 1, logs
        text = open(logs[0], encoding='utf-8').read()
        with open(os.path.join(HERE, name + '.txt'), 'w', encoding='utf-8') as f:
            f.write(os.path.basename(logs[0]) + '\n')        # first line: the log file name the reference chose
            f.write(strip_timing(text))
        print('golden', name, os.path.basename(logs[0]))
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    main()

"""Multi-GPU host logic on CPU: the sentence partition and the world_size-2 result gather (gloo).

The data path has no collective; what is covered here is that two ranks agree on the
partition without talking, that every sentence is decoded exactly once, and that the gathered n-best
lists come back in input order and equal a single-process decode.  The CPU oracle stands in for the GPU
decoder (it exposes the same decode_batch contract through a thin adapter)."""
import os
import socket
import sys

import pytest

from jlm_b200 import shard, synth

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_everything_once_and_balances():
    lengths = [20, 3, 41, 7, 7, 25, 1, 30, 22, 22, 5]
    for world in (1, 2, 3, 8):
        shards = shard.partition(lengths, world)
        assert len(shards) == world
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lengths)))
        assert all(s == sorted(s) for s in shards)
        loads = [sum(lengths[i] + 1 for i in s) for s in shards]
        if world <= len(lengths):
            assert max(loads) - min(loads) <= max(lengths) + 1
    assert shard.partition([], 2) == [[], []]
    assert shard.partition(lengths, 2) == shard.partition(list(lengths), 2)       # deterministic
    with pytest.raises(ValueError):
        shard.partition(lengths, 0)


class _OracleBatchDecoder(object):
    """decode_batch adapter over the CPU oracle (test stand-in for jlm_b200.Decoder)."""

    def __init__(self):
        from oracle import jlm_oracle as O
        cfg = synth.make_config(300, 32, 16, 'tied')
        weights = synth.make_weights(cfg, seed=3)
        lexicon, reading = synth.make_lexicon(300, seed=3)
        self.dec = O.OracleDecoder(cfg, weights, lexicon, reading)
        self.sentences = synth.make_sentences(lexicon, 9, min_len=6, seed=4, vocab_size=300)
        self.calls = 0

    def decode_batch(self, texts, **kw):
        self.calls += len(texts)
        return [self.dec.decode(t, **kw) for t in texts]


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        d = _OracleBatchDecoder()
        out = shard.decode_sharded(d, d.sentences, rank=rank, world_size=world, gather=True, topN=3, beam_width=4)
        q.put((rank, d.calls, out))
    finally:
        dist.destroy_process_group()


def test_decode_sharded_world2_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _OracleBatchDecoder()
    want = ref.decode_batch(ref.sentences, topN=3, beam_width=4)
    calls = 0
    for rank, n_calls, out in got:
        assert out == want, rank                      # every rank holds the full list, in input order
        calls += n_calls
    assert calls == len(ref.sentences)                # each sentence decoded exactly once across ranks
    # without gather a rank only fills its own shard
    part = shard.decode_sharded(ref, ref.sentences, rank=1, world_size=2, gather=False, topN=3, beam_width=4)
    mine = shard.partition([len(t) for t in ref.sentences], 2)[1]
    assert [i for i, r in enumerate(part) if r is not None] == mine

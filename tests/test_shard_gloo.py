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


class _OraclePackedDecoder(_OracleBatchDecoder):
    """The same stand-in with the array form of the n-best block (Decoder.decode_batch_arrays /
    words_from_arrays contract): paths as lexicon entry ids, -2 + start frame for '<unk>' nodes."""

    def __init__(self):
        super().__init__()
        import numpy as np
        self.np = np
        lexicon = self.dec.lexicon if hasattr(self.dec, 'lexicon') else None
        if lexicon is None:
            lexicon, _ = synth.make_lexicon(300, seed=3)
        self.entry_words = [w for w, _ in lexicon]
        self.entry_of = {w: i for i, w in enumerate(self.entry_words)}

    def decode_batch_arrays(self, texts, topN=10, beam_width=10, **kw):
        np = self.np
        res = self.decode_batch(texts, topN=topN, beam_width=beam_width)
        top = min(topN, beam_width)
        L = max(len(t) for t in texts) + 1
        a = {'scores': np.full((len(texts), top), np.inf), 'n_paths': np.zeros(len(texts), dtype=np.int32),
             'path_len': np.zeros((len(texts), top), dtype=np.int32),
             'path_entry': np.zeros((len(texts), top, L), dtype=np.int32),
             'path_start': np.zeros((len(texts), top, L), dtype=np.int32)}
        for s, r in enumerate(res):
            a['n_paths'][s] = len(r)
            for k, (score, words) in enumerate(r):
                a['scores'][s, k] = score
                a['path_len'][s, k] = len(words)
                pos = 0
                for q, w in enumerate(words):
                    known = w in self.entry_of
                    a['path_entry'][s, k, q] = self.entry_of[w] if known else -2
                    a['path_start'][s, k, q] = pos
                    pos += len(w.split('/')[1]) if known else 1
        return a

    def words_from_arrays(self, texts, a, topN=None):
        out = []
        for s, text in enumerate(texts):
            res = []
            for k in range(int(a['n_paths'][s])):
                n = int(a['path_len'][s, k])
                res.append((float(a['scores'][s, k]),
                            [self.entry_words[e] if e >= 0 else text[st]
                             for e, st in zip(a['path_entry'][s, k, :n].tolist(), a['path_start'][s, k, :n].tolist())]))
            out.append(res[:topN] if topN is not None else res)
        return out


def _worker(rank, world, port, q, packed=False):
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        d = _OraclePackedDecoder() if packed else _OracleBatchDecoder()
        out = shard.decode_sharded(d, d.sentences, rank=rank, world_size=world, gather=True, topN=3, beam_width=4)
        q.put((rank, d.calls, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('packed', [False, True])
def test_decode_sharded_world2_gloo(packed):
    """packed=True: the n-best block travels as two fixed-shape tensors (all_gather_into_tensor) and every rank
    rebuilds the word lists; packed=False: decoders without the array form fall back to all_gather_object."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, packed)) for r in range(2)]
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


def test_pack_unpack_roundtrip_and_thread_budget():
    import numpy as np
    rng = np.random.default_rng(0)
    a = {'scores': rng.normal(size=(3, 2)), 'n_paths': np.array([2, 1, 0], dtype=np.int32),
         'path_len': rng.integers(0, 5, size=(3, 2)).astype(np.int32),
         'path_entry': rng.integers(-2, 50, size=(3, 2, 5)).astype(np.int32),
         'path_start': rng.integers(0, 5, size=(3, 2, 5)).astype(np.int32)}
    f, i = shard.pack_arrays(a, 4, 3, 7)           # padded to 4 rows, top 3, max_len 7
    assert f.shape == (4, 3) and i.shape == (4, 1 + 3 + 2 * 3 * 7)
    b = shard.unpack_arrays(f, i, 3, 3, 7)
    assert np.array_equal(b['scores'][:, :2], a['scores']) and np.all(np.isinf(b['scores'][:, 2]))
    assert np.array_equal(b['n_paths'], a['n_paths']) and np.array_equal(b['path_len'][:, :2], a['path_len'])
    assert np.array_equal(b['path_entry'][:, :2, :5], a['path_entry'])
    assert np.array_equal(b['path_start'][:, :2, :5], a['path_start'])
    f0, i0 = shard.pack_arrays(None, 2, 3, 7)      # a rank with an empty shard still contributes its block
    assert f0.shape == (2, 3) and not i0.any()
    assert shard.host_threads_per_rank(1) == (os.cpu_count() or 1)
    assert shard.host_threads_per_rank(10 ** 6) == 1


def test_decode_sharded_as_arrays_single_rank():
    """as_arrays=True returns the whole n-best block in input order; world_size 1 needs no process group."""
    d = _OraclePackedDecoder()
    want = d.decode_batch(d.sentences, topN=3, beam_width=4)
    arr = shard.decode_sharded(d, d.sentences, rank=0, world_size=1, gather=True, as_arrays=True, topN=3, beam_width=4)
    assert arr['scores'].shape == (len(d.sentences), 3)
    assert d.words_from_arrays(d.sentences, arr, 3) == want
    with pytest.raises(ValueError):
        shard.decode_sharded(_OracleBatchDecoder(), d.sentences, rank=0, world_size=1, as_arrays=True, topN=3, beam_width=4)

"""Data-parallel sharding of independent sentences across GPUs.

The path has no exchange step: every sentence is decoded on exactly one GPU against a replica of the
model, so the only multi-GPU logic is (a) a deterministic, length-balanced partition of the input list
and (b) putting the per-rank n-best lists back in input order.  One process per GPU (torchrun);
``torch.distributed`` is used for the result gather only - never inside the decode.

The gather exchanges the n-best block in its ARRAY form (scores float64, paths as int32 lexicon entries +
start frames, what ``Decoder.decode_batch_arrays`` returns): two fixed-shape ``all_gather_into_tensor``
calls (NCCL over NVLink on a GPU box, gloo in the CPU tests) instead of pickled Python lists; every rank
then rebuilds the word strings from its own copy of the lexicon.
"""
import os

import numpy as np


def partition(lengths, world_size):
    """Splits sentence indices over ``world_size`` ranks, balancing the lock-step cost sum(T+1).

    Longest-first greedy onto the least-loaded rank (ties -> lowest rank), so every rank computes the
    same partition from the same input without communicating.  Each shard keeps input order."""
    if world_size < 1:
        raise ValueError('world_size must be >= 1')
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += lengths[i] + 1
    return [sorted(s) for s in shards]


def env_rank_world():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def host_threads_per_rank(world_size):
    """Host threads one rank may use for lattice / plan building when ``world_size`` ranks share the box."""
    return max(1, (os.cpu_count() or 1) // max(1, world_size))


def pack_arrays(arrays, n_rows, top, max_len):
    """One float64 and one int32 buffer of fixed shape [n_rows, ...] from a shard's n-best arrays (padded)."""
    n = arrays['scores'].shape[0] if arrays is not None else 0
    f = np.full((n_rows, top), np.inf)
    i = np.zeros((n_rows, 1 + top + 2 * top * max_len), dtype=np.int32)
    if n:
        t, L = arrays['scores'].shape[1], arrays['path_entry'].shape[2]
        f[:n, :t] = arrays['scores']
        i[:n, 0] = arrays['n_paths']
        i[:n, 1:1 + t] = arrays['path_len']
        e = np.zeros((n, top, max_len), dtype=np.int32)
        s = np.zeros((n, top, max_len), dtype=np.int32)
        e[:, :t, :L] = arrays['path_entry']
        s[:, :t, :L] = arrays['path_start']
        i[:n, 1 + top:1 + top + top * max_len] = e.reshape(n, -1)
        i[:n, 1 + top + top * max_len:] = s.reshape(n, -1)
    return f, i


def unpack_arrays(f, i, n, top, max_len):
    """Inverse of pack_arrays for the first ``n`` rows."""
    return {'scores': f[:n], 'n_paths': i[:n, 0], 'path_len': i[:n, 1:1 + top],
            'path_entry': i[:n, 1 + top:1 + top + top * max_len].reshape(n, top, max_len),
            'path_start': i[:n, 1 + top + top * max_len:].reshape(n, top, max_len)}


def decode_sharded(decoder, texts, rank=None, world_size=None, gather=True, as_arrays=False, **decode_kwargs):
    """``decoder.decode_batch`` over this rank's shard of ``texts``.

    With ``gather=True`` and an initialised process group every rank returns the full list in input order.
    Decoders that offer ``decode_batch_arrays`` (the GPU decoders) exchange packed arrays with
    ``all_gather_into_tensor``; others fall back to ``all_gather_object``.  With ``gather=False`` a list with
    ``None`` for sentences owned by other ranks is returned.  ``as_arrays=True`` (GPU decoders) returns the n-best
    block of the whole input, in input order, in its array form (``decode_batch_arrays``) instead of Python word
    lists - ``decoder.words_from_arrays(texts, arrays)`` turns any part of it into words later."""
    if rank is None or world_size is None:
        rank, world_size = env_rank_world()
    texts = list(texts)
    shards = partition([len(t) for t in texts], world_size)
    mine = shards[rank]
    out = [None] * len(texts)
    packed = hasattr(decoder, 'decode_batch_arrays') and gather and (world_size > 1 or as_arrays)
    if as_arrays and not packed:
        raise ValueError('as_arrays=True needs gather=True and a decoder with decode_batch_arrays')
    if not packed:
        local = decoder.decode_batch([texts[i] for i in mine], **decode_kwargs) if mine else []
        for i, res in zip(mine, local):
            out[i] = res
        if gather and world_size > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError('decode_sharded(gather=True) needs an initialised process group')
            parts = [None] * world_size
            dist.all_gather_object(parts, list(zip(mine, local)))
            for part in parts:
                for i, res in part:
                    out[i] = res
        return out

    topN = int(decode_kwargs.get('topN', 10))
    beam = decode_kwargs.get('beam_width', 10)
    top = max(1, topN if beam is None else min(topN, int(beam)))
    max_len = max((len(t) for t in texts), default=0) + 1
    n_rows = max(len(s) for s in shards)
    arrays = decoder.decode_batch_arrays([texts[i] for i in mine], **decode_kwargs) if mine else None
    f, i32 = pack_arrays(arrays, n_rows, top, max_len)
    if world_size > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('decode_sharded(gather=True) needs an initialised process group')
        dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
        tf, ti = torch.from_numpy(f).to(dev), torch.from_numpy(i32).to(dev)
        # output = the ranks' blocks concatenated along dim 0 (the layout both NCCL and gloo accept)
        gf = torch.empty((world_size * tf.shape[0], tf.shape[1]), dtype=tf.dtype, device=dev)
        gi = torch.empty((world_size * ti.shape[0], ti.shape[1]), dtype=ti.dtype, device=dev)
        dist.all_gather_into_tensor(gf, tf)
        dist.all_gather_into_tensor(gi, ti)
        gf = gf.cpu().numpy().reshape(world_size, n_rows, -1)
        gi = gi.cpu().numpy().reshape(world_size, n_rows, -1)
    else:
        gf, gi = f[None], i32[None]
    if as_arrays:
        # scatter the ranks' rows back to input order
        full_f = np.empty((len(texts), gf.shape[2]), dtype=gf.dtype)
        full_i = np.empty((len(texts), gi.shape[2]), dtype=gi.dtype)
        for r, idx in enumerate(shards):
            if idx:
                full_f[idx] = gf[r, :len(idx)]
                full_i[idx] = gi[r, :len(idx)]
        return unpack_arrays(full_f, full_i, len(texts), top, max_len)
    for r, idx in enumerate(shards):
        if not idx:
            continue
        words = decoder.words_from_arrays([texts[k] for k in idx], unpack_arrays(gf[r], gi[r], len(idx), top, max_len), topN)
        for k, res in zip(idx, words):
            out[k] = res
    return out

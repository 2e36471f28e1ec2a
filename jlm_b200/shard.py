"""Data-parallel sharding of independent sentences across GPUs.

The path has no exchange step: every sentence is decoded on exactly one GPU against a replica of the
model, so the only multi-GPU logic is (a) a deterministic, length-balanced partition of the input list
and (b) putting the per-rank n-best lists back in input order.  One process per GPU (torchrun);
``torch.distributed`` is used for the result gather only - never inside the decode.
"""
import os


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


def decode_sharded(decoder, texts, rank=None, world_size=None, gather=True, **decode_kwargs):
    """``decoder.decode_batch`` over this rank's shard of ``texts``.

    With ``gather=True`` and an initialised process group the per-rank results are exchanged with
    ``all_gather_object`` and every rank returns the full list in input order; otherwise a list with
    ``None`` for sentences owned by other ranks is returned."""
    if rank is None or world_size is None:
        rank, world_size = env_rank_world()
    texts = list(texts)
    shards = partition([len(t) for t in texts], world_size)
    mine = shards[rank]
    local = decoder.decode_batch([texts[i] for i in mine], **decode_kwargs) if mine else []
    out = [None] * len(texts)
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

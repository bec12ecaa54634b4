"""Utterance sharding across GPUs.

Utterances never interact on the synthesis path (every op is per batch row,
/root/reference/networks.py:140-153, synthesize.py:54), so the batch is split into
contiguous shards, one per rank, with weights replicated.  The only communication is
ONE gather of the finished spectrograms; the reference has no counterpart (it is a
single-process program).  Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.

Training (BASELINE config 5) is plain data parallelism: every rank runs `Engine.train_step(..., apply=False)`
on its own 32 utterances, the flat gradient arena (`Engine.train_grads()`) is averaged over the ranks
(`allreduce_mean_`), then every rank applies the identical Adam update (`Engine.train_apply`).
"""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous split of `total` utterances: rank g gets [lo, hi); sizes differ by <= 1."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_spectrograms(local, total, dst=0, group=None):
    """Gather per-rank (b_g, T, F) tensors into the (total, T, F) tensor on `dst`
    (None elsewhere).  Shards may be ragged by one utterance; rows keep global order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    out = None
    if rank == dst:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    ops = []
    if rank == dst:
        for g in range(world):
            lo, hi = shard_bounds(total, g, world)
            if g == dst:
                out[lo:hi].copy_(local)
            elif hi > lo:
                ops.append(dist.P2POp(dist.irecv, out[lo:hi], g, group))
    else:
        lo, hi = shard_bounds(total, rank, world)
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, local.contiguous(), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out


def allreduce_mean_(flat, group=None):
    """In-place average of a flat gradient buffer over the ranks (sum all-reduce, then 1/world): the gradient of the mean
    loss over the global batch when every rank holds the same number of utterances.  Returns `flat`."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    return flat

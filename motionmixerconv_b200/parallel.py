"""Batch-sharded data parallelism: host-side logic (device-agnostic, so the world_size-2 gloo tests can run it on CPU).

One process per GPU; rank r owns sequences [r*B/W, (r+1)*B/W) of the global batch; every rank holds all parameters,
gradients travel as ONE flat fp32 bucket with a single all-reduce(SUM) per step; the 1/W of the mean is folded into the
fused Adam (``grad_scale``).  There is no other collective on the path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

ALIGN = 4   # floats: every parameter view starts 16-byte aligned in the flat buffers


def flat_offsets(numels):
    """Offsets (in floats) of each tensor inside the flat bucket, and the bucket length."""
    offs, o = [], 0
    for n in numels:
        offs.append(o)
        o += (int(n) + ALIGN - 1) // ALIGN * ALIGN
    return offs, o


def shard_bounds(batch, rank, world):
    """[lo, hi) of rank's shard; the global batch must divide evenly (equal shards keep mean-of-means == global mean)."""
    if batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (batch, world))
    per = batch // world
    return rank * per, (rank + 1) * per


def shard(t, rank, world):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_bucket(flat_grads: torch.Tensor, group=None):
    """SUM the flat gradient bucket over the group (NCCL on GPUs, gloo in the CPU tests).  Returns grad_scale = 1/W,
    the factor the optimiser applies."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world

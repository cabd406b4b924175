"""Multi-GPU helpers for the rasterizer.

The batch / viewpoint axis shards embarrassingly: every kernel treats the batch index as an independent outer
dimension, so one process per GPU renders its own contiguous slice and no data-path collective is needed
(`shard_range`).  The only exchange the path ever has is the gradient of a mesh SHARED by all viewpoints
(`Mesh.get_batch` broadcasts one mesh, mesh.py:29-34 of the reference): each rank reduces its own views locally
(autograd sums over the expanded batch axis) and the per-rank sums of `vertices.grad` / `textures.grad` are combined
with one sum-all-reduce each (`allreduce_shared_grads`; NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` batch items / viewpoints owned by `rank`."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def allreduce_shared_grads(params, group=None, async_op=False):
    """Sum-all-reduce the `.grad` of parameters shared by all ranks (in place).  Large tensors are reduced as they
    are (no flattening copy); with async_op=True the work handles are returned so that the caller can overlap the
    texture-gradient reduction with other work and `wait()` later."""
    import torch.distributed as dist
    works = []
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return works
    for p in params:
        g = p.grad if isinstance(p, torch.Tensor) and p.grad is not None else None
        if g is None:
            continue
        w = dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works

"""Replica sharding across ranks (one process per GPU).  Replicas are independent, so the
simulation path has no collective; the learner all-reduces (SUM) its flat gradient once per update
after each rank scaled its local sum by 1 / (n_step * total_replicas)."""
from __future__ import annotations

import numpy as np


def shard_replicas(rank: int, world: int, replicas_per_rank: int, seed0: int):
    """Global replica ids and episode seeds of one rank: rank k owns [k*R, (k+1)*R)."""
    replica0 = rank * replicas_per_rank
    ids = np.arange(replicas_per_rank, dtype=np.int64) + replica0
    seeds = (ids + seed0).astype(np.uint64)
    return replica0, ids, seeds


def grad_scale(n_step: int, world: int, replicas_per_rank: int) -> float:
    return 1.0 / (n_step * world * replicas_per_rank)


def allreduce_sum_(flat_grad, group=None):
    import torch.distributed as dist
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad

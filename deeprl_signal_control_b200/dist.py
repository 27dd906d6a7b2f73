"""Replica sharding across ranks (one process per GPU).  Replicas are independent, so the
simulation path has no collective; the learner all-reduces (SUM) its flat gradient once per update
after each rank scaled its local sum by 1 / (n_step * total_replicas).

Used by the product path: `BatchedTrainer.start_episode` (episode_seeds), `BatchedA2C.backward` (grad_scale,
allreduce_sum_) and bench.py (shard_replicas)."""
from __future__ import annotations

import numpy as np


def shard_replicas(rank: int, world: int, replicas_per_rank: int, seed0: int):
    """Global replica ids and first-episode seeds of one rank: rank k owns [k*R, (k+1)*R)."""
    replica0 = rank * replicas_per_rank
    ids = np.arange(replicas_per_rank, dtype=np.int64) + replica0
    seeds = episode_seeds(seed0, 0, replica0, replicas_per_rank, world * replicas_per_rank)
    return replica0, ids, seeds


def episode_seeds(seed0: int, episode: int, replica0: int, n_local: int, total_replicas: int) -> np.ndarray:
    """Seeds of this rank's replicas for its `episode`-th episode.  The reference re-seeds SUMO with `seed += 1` per
    episode (envs/env.py:560); with R_total lock-stepped replicas the counter advances by R_total per episode, so that
    the seed sets of all (rank, episode) pairs are disjoint: seed = seed0 + episode * R_total + global replica id."""
    ids = np.arange(n_local, dtype=np.int64) + int(replica0)
    return (ids + int(seed0) + int(episode) * int(total_replicas)).astype(np.uint64)


def grad_scale(n_step: int, world: int, replicas_per_rank: int) -> float:
    return 1.0 / (n_step * world * replicas_per_rank)


def allreduce_sum_(flat_grad, group=None):
    import torch.distributed as dist
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad

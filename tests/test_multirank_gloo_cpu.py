"""world_size-2 gloo test (CPU) of the multi-GPU host logic: replica sharding and the gradient
convention (local sum scaled by 1/(n_step*R_total), all-reduce SUM == full-batch mean gradient)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    from deeprl_signal_control_b200.dist import allreduce_sum_, grad_scale, shard_replicas
    from oracle.learner_ref import a2c_loss
    n_s, n_w, n_f = [32, 42], [6, 6], [8, 12]
    off = np.array([0, 32, 74], np.int32)
    lay = PolicyLayout(n_s, [5, 5], n_w, n_f, off, 74, fw=16, ft=8, ff=8, h=64, max_na=5)
    T, R = 3, 4                                   # replicas per rank
    replica0, ids, seeds = shard_replicas(rank, world, R, seed0=12)
    assert replica0 == rank * R and list(seeds) == [12 + rank * R + r for r in range(R)]
    g = torch.Generator().manual_seed(0)          # identical full batch on every rank; each takes its shard
    Rt = world * R
    obs = torch.rand(T, Rt, 74, generator=g, dtype=torch.float64)
    acts = torch.randint(0, 5, (T, Rt, 2), generator=g)
    Rs = torch.randn(T, Rt, 2, generator=g, dtype=torch.float64)
    Adv = torch.randn(T, Rt, 2, generator=g, dtype=torch.float64)
    P0 = torch.from_numpy(lay.init_params(1).astype(np.float64))
    zeros = [torch.zeros(Rt, 64, dtype=torch.float64) for _ in range(lay.U)]

    def grad(sl):
        P = P0.clone().requires_grad_(True)
        n = sl.stop - sl.start
        loss, _ = a2c_loss(P, lay, obs[:, sl], acts[:, sl], Rs[:, sl], Adv[:, sl], [0.0] * T,
                           [z[sl] for z in zeros], [z[sl] for z in zeros], 0.5, 0.01)
        # a2c_loss averages over (t, local replicas): turn it into the local SUM scaled by the global factor
        (loss * (T * n) * grad_scale(T, world, R)).backward()
        return P.grad

    local = grad(slice(rank * R, (rank + 1) * R))
    allreduce_sum_(local)
    full = grad(slice(0, Rt)) / ((T * Rt) * grad_scale(T, world, R))   # plain mean gradient of the full batch
    assert torch.allclose(local, full, rtol=1e-9, atol=1e-12)
    open(os.path.join(out_dir, "ok%d" % rank), "w").write("1")
    dist.destroy_process_group()


def test_gradient_allreduce_convention_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_episode_seed_sets_are_disjoint_across_ranks_and_episodes():
    """Rank k, episode e, replica r plays seed0 + e*R_total + k*R + r: no (rank, episode) pair replays the demand of
    another one (the 1/(T*R_total) gradient scaling assumes R_total decorrelated trajectories per update)."""
    from deeprl_signal_control_b200.dist import episode_seeds, shard_replicas
    world, R, seed0 = 4, 8, 12
    seen = set()
    for rank in range(world):
        replica0, ids, s0 = shard_replicas(rank, world, R, seed0)
        assert list(s0) == list(episode_seeds(seed0, 0, replica0, R, world * R))
        for ep in range(5):
            s = set(int(x) for x in episode_seeds(seed0, ep, replica0, R, world * R))
            assert len(s) == R and not (s & seen)
            seen |= s
    assert seen == set(range(seed0, seed0 + 5 * world * R))      # and contiguous: the reference's seed += 1 per episode

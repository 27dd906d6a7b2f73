"""CPU leg of the bench's reference arm for the LEARNER (TEST / BENCH INFRASTRUCTURE ONLY — never imported by the
product package): the same work the GPU arm does per control step, on the host cores, in torch CPU fp32:

  * policy forward of every agent (two networks each) for a batch of replicas, LSTM state carried  (reference
    agents/policies.py:125-136 runs this as 25-50 batch-1 sess.run calls per env step);
  * one n-step A2C update: loss of agents/policies.py:41-52 back-propagated through the n_step unroll by autograd,
    per-agent clip_by_global_norm + TF1 RMSProp (the restatement pinned in tests/test_learner_reference_golden_cpu.py).

It re-uses the float-agnostic restatement in oracle/learner_ref.py (unit_forward / a2c_loss / clip_rmsprop) with
float32 tensors; TensorFlow 1.12 itself is absent from the image, hence kind = "port".
"""
import time

import numpy as np
import torch

from .learner_ref import clip_rmsprop, unit_forward


class CpuA2C:
    def __init__(self, lay, n_replicas, n_step, seed=0, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.lay, self.R, self.T = lay, int(n_replicas), int(n_step)
        self.P = torch.from_numpy(lay.init_params(seed))
        self.MS = np.ones(lay.n_params, np.float32)
        self.c = [torch.zeros(self.R, lay.h) for _ in range(lay.U)]
        self.h = [torch.zeros(self.R, lay.h) for _ in range(lay.U)]

    @torch.no_grad()
    def forward(self, obs, done):
        """obs float32 [R, n_obs] -> (pi [R, A, max_na], val [R, A]); advances the LSTM states."""
        lay = self.lay
        v = lay.views(self.P)
        pi = torch.zeros(self.R, lay.A, lay.max_na)
        val = torch.zeros(self.R, lay.A)
        o = obs[None]
        for a in range(lay.A):
            p, _, self.c[2 * a], self.h[2 * a] = unit_forward(v, lay, 2 * a, o, [float(done)], self.c[2 * a], self.h[2 * a])
            w, _, self.c[2 * a + 1], self.h[2 * a + 1] = unit_forward(v, lay, 2 * a + 1, o, [float(done)],
                                                                      self.c[2 * a + 1], self.h[2 * a + 1])
            pi[:, a, :p.shape[-1]] = p[0]
            val[:, a] = w[0]
        return pi, val

    def update(self, obs, acts, Rs, Advs, dones, lr=5e-4, beta=0.01, v_coef=0.5, max_norm=40.0):
        """One n-step update; agents are back-propagated one at a time (the reference also trains its 25 policies
        sequentially, agents/models.py:177-183), which bounds the autograd memory to one agent's unroll."""
        lay = self.lay
        R = obs.shape[1]
        P = self.P.clone().requires_grad_(True)
        v = lay.views(P)
        z = torch.zeros(R, lay.h)
        for a in range(lay.A):
            pi, _, _, _ = unit_forward(v, lay, 2 * a, obs, dones, z, z)
            val, _, _, _ = unit_forward(v, lay, 2 * a + 1, obs, dones, z, z)
            log_pi = torch.log(torch.clamp(pi, 1e-10, 1.0))
            ent = -(pi * log_pi).sum(-1)
            lp_a = torch.gather(log_pi, -1, acts[..., a].long().unsqueeze(-1)).squeeze(-1)
            loss = -(lp_a * Advs[..., a]).mean() + ((Rs[..., a] - val) ** 2).mean() * 0.5 * v_coef - ent.mean() * beta
            loss.backward()
        Pn, MS, _ = clip_rmsprop(self.P.numpy(), P.grad.numpy(), self.MS, lay.agent_of, max_norm, lr, 0.99, 1e-5, lay.A)
        self.P, self.MS = torch.from_numpy(Pn.astype(np.float32)), MS.astype(np.float32)


def time_learner(lay, R_fwd, n_fwd, R_upd, n_step, threads, seed=0):
    """Seconds per policy forward of R_fwd replicas, and seconds of one n_step update of R_upd replicas."""
    g = torch.Generator().manual_seed(seed)
    m = CpuA2C(lay, R_fwd, n_step, seed=seed, threads=threads)
    obs = torch.rand(R_fwd, lay.n_obs, generator=g) * 2
    m.forward(obs, True)
    t0 = time.perf_counter()
    for _ in range(n_fwd):
        m.forward(obs, False)
    t_fwd = (time.perf_counter() - t0) / n_fwd
    T = n_step
    obs_u = torch.rand(T, R_upd, lay.n_obs, generator=g) * 2
    acts = torch.stack([torch.randint(0, int(n), (T, R_upd), generator=g) for n in lay.n_a], -1)
    Rs = torch.randn(T, R_upd, lay.A, generator=g)
    Adv = torch.randn(T, R_upd, lay.A, generator=g)
    t0 = time.perf_counter()
    m.update(obs_u, acts, Rs, Adv, [0.0] * T)
    t_upd = time.perf_counter() - t0
    return t_fwd, t_upd

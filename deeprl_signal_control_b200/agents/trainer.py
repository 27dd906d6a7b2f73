"""Device-resident training loop: the batched counterpart of reference utils.py:Trainer.explore/run.

Per control step (utils.py:146-165): policy/value forward -> fingerprint update (MA2C) -> action
sampling -> env.step -> add_transition; per n_step: bootstrap value, backward (utils.py:186-190,
288-291); per episode: env.reset(), model.reset(), pre-decision done = True (utils.py:277-281).
Everything stays on the GPU: the simulator writes the next observation straight into the
learner's rollout slot and reads actions / fingerprints from the learner's buffers.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .. import dist as _dist
from ..sim import BatchedSim
from .learner import BatchedA2C


class BatchedTrainer:
    def __init__(self, sim: BatchedSim, model: BatchedA2C, agent: str, lr, beta,
                 seed0: int = 12, replica0: int = 0):
        """lr / beta: floats (the 'constant' schedules of every shipped A2C config) or objects with the reference's
        `Scheduler.get(n_step)` (agents/utils.py:268-281, agents/models.py:175-176); a schedule advances by n_step per
        update exactly as in the reference (its unit is control steps of ONE environment)."""
        self.sim, self.model, self.agent = sim, model, agent
        self.lr, self.beta = lr, beta
        self.seed0, self.replica0 = int(seed0), int(replica0)
        self.total_replicas = int(getattr(model, "total_replicas", sim.R))
        self.episode = 0
        self.T_episode = int(np.ceil(sim.params.episode_length_sec / sim.params.control_interval_sec))
        assert self.T_episode % model.T == 0                      # utils.py:121
        self.step_in_episode = 0
        self.done = True
        self.episode_rewards = []
        self._rew_acc = torch.zeros(sim.R, device=sim.device)
        self.n_updates = 0
        self.n_env_steps = 0
        self._uniform_fp = None
        self.sim_events = None        # list of (start, end) CUDA events around tsc_step when timing is on
        self.update_events = None     # same around update() (bootstrap forward + backward)
        self.start_episode()

    def start_episode(self):
        sim, m = self.sim, self.model
        # envs/env.py:560 (seed += 1 per episode and environment): disjoint over all (rank, episode) pairs
        seeds = _dist.episode_seeds(self.seed0, self.episode, self.replica0, sim.R, max(self.total_replicas, sim.R))
        self.episode += 1
        sim.reset(seeds)
        sim.set_train_mode(True)
        m.reset()
        fp = None
        if self.agent == 'ma2c':
            if self._uniform_fp is None:
                n = sim.net
                u = torch.zeros(sim.R, n.n_nodes, n.max_na, device=sim.device)
                for i, na in enumerate(n.n_a_ls):
                    u[:, i, :na] = 1.0 / na                       # envs/env.py:263-269
                self._uniform_fp = u
            fp = self._uniform_fp
        assert m.t == 0
        sim.observe(fp, obs_out=m.obs_slot(0))
        self.step_in_episode = 0
        self.done = True
        self._rew_acc.zero_()

    def control_step(self):
        """One control step of all replicas (utils.py:146-165)."""
        sim, m = self.sim, self.model
        t = m.t
        pi, val, act = m.forward(m.obs_slot(t), self.done)
        fp = pi if self.agent == 'ma2c' else None                 # env.update_fingerprint(policy)
        if self.sim_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _, reward, greward, _ = sim.step(act, fp, obs_out=m.obs_hist[t + 1])
        if self.sim_events is not None:
            e1.record()
            self.sim_events.append((e0, e1))
        self.step_in_episode += 1
        new_done = self.step_in_episode >= self.T_episode         # lock-step: envs/env.py:577-579
        m.add_transition(reward, self.done, new_done)
        self._rew_acc.add_(greward)
        self.done = new_done
        self.n_env_steps += 1
        if m.t == m.T:
            self.update()

    def control_step_host(self):
        """Same step with the environment driven through the HOST-buffer C-ABI call (tsc_step_host),
        i.e. the way a reference-style caller holds numpy arrays: actions / fingerprints D2H, env step
        (H2D + kernel + D2H inside), observation / reward H2D into the learner."""
        sim, m = self.sim, self.model
        t = m.t
        pi, val, act = m.forward(m.obs_slot(t), self.done)
        if not hasattr(self, '_pin_act'):       # page-locked staging for the D2H of actions / fingerprints
            self._pin_act = torch.zeros_like(act, device='cpu').pin_memory()
            self._pin_pi = torch.zeros_like(pi, device='cpu').pin_memory()
        self._pin_act.copy_(act, non_blocking=True)
        if self.agent == 'ma2c':
            self._pin_pi.copy_(pi, non_blocking=True)
        torch.cuda.current_stream(sim.device).synchronize()
        act_h = self._pin_act.numpy()
        fp_h = self._pin_pi.numpy() if self.agent == 'ma2c' else None
        obs_h, rew_h, grew_h, _ = sim.step_host(act_h, fp_h)
        m.obs_hist[t + 1].copy_(torch.from_numpy(obs_h), non_blocking=True)
        reward = torch.from_numpy(rew_h).to(sim.device, non_blocking=True)
        self.step_in_episode += 1
        new_done = self.step_in_episode >= self.T_episode
        m.add_transition(reward, self.done, new_done)
        self._rew_acc.add_(torch.from_numpy(grew_h).to(sim.device, non_blocking=True))
        self.done = new_done
        self.n_env_steps += 1
        if m.t == m.T:
            self.update()

    def control_step_host_pipelined(self, n_parts: int = 2):
        """control_step_host() with the replicas split into `n_parts` ranges, one CUDA stream each: while range k sits on
        the PCIe link (tsc_step_host_range: actions H2D, kernel, observations D2H; then observations H2D into the
        learner), the other ranges run their policy forward / simulator kernels.  Same results as control_step_host():
        ranges are independent and the sampling RNG is keyed by the absolute replica.  The host still receives every
        range's observations / rewards in page-locked numpy buffers before the learner consumes them."""
        sim, m = self.sim, self.model
        R, A = sim.R, m.lay.A
        if not hasattr(self, '_pp'):
            n_parts = max(1, min(n_parts, R))
            bounds = [R * k // n_parts for k in range(n_parts + 1)]
            pin = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype).pin_memory()
            self._pp = dict(parts=[(bounds[k], bounds[k + 1] - bounds[k]) for k in range(n_parts)],
                            streams=[torch.cuda.Stream(device=sim.device) for _ in range(n_parts)],
                            ev=[torch.cuda.Event() for _ in range(n_parts)], primed=False,
                            act=pin(R, A, dtype=torch.int32), pi=pin(R, A, m.lay.max_na), obs=pin(R, sim.net.n_obs),
                            rew=pin(R, A), grew=pin(R), done=pin(R, dtype=torch.uint8))
        pp = self._pp
        ma2c = self.agent == 'ma2c'

        def issue_forward(k, t, nf, done):
            r0, n = pp['parts'][k]
            pi, val, act = m.forward_range(r0, n, done, t, nf)
            pp['act'][r0:r0 + n].copy_(act, non_blocking=True)
            if ma2c:
                pp['pi'][r0:r0 + n].copy_(pi, non_blocking=True)
            pp['ev'][k].record()

        if not pp['primed']:
            cur = torch.cuda.current_stream(sim.device)
            for k, st in enumerate(pp['streams']):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    issue_forward(k, m.t, m.n_forward, self.done)
            m.end_forward_ranges()
            pp['primed'] = True
        t = m.t
        self.step_in_episode += 1
        new_done = self.step_in_episode >= self.T_episode
        boundary = (t + 1 == m.T)                       # an update (and possibly an episode end) follows this step
        npv = lambda x, r0, n: x[r0:r0 + n].numpy()
        for k, st in enumerate(pp['streams']):       # every range's env step is enqueued as soon as its actions are here
            r0, n = pp['parts'][k]
            pp['ev'][k].synchronize()                   # the host holds this range's actions / fingerprints
            with torch.cuda.stream(st):
                sim.step_host_range(r0, n, npv(pp['act'], r0, n), npv(pp['pi'], r0, n) if ma2c else None,
                                    npv(pp['obs'], r0, n), npv(pp['rew'], r0, n), npv(pp['grew'], r0, n),
                                    npv(pp['done'], r0, n), sync=False)
        for k, st in enumerate(pp['streams']):
            r0, n = pp['parts'][k]
            st.synchronize()                            # the host holds this range's observations / rewards ...
            with torch.cuda.stream(st):
                # ... and hands them to the learner
                m.obs_hist[t + 1, r0:r0 + n].copy_(pp['obs'][r0:r0 + n], non_blocking=True)
                m.add_transition_range(r0, n, pp['rew'][r0:r0 + n].to(sim.device, non_blocking=True))
                self._rew_acc[r0:r0 + n].add_(pp['grew'][r0:r0 + n].to(sim.device, non_blocking=True))
                if not boundary:
                    issue_forward(k, t + 1, m.n_forward, False)
        m.end_transition_ranges(self.done, new_done)
        self.done = new_done
        self.n_env_steps += 1
        if not boundary:
            m.end_forward_ranges()
        else:
            cur = torch.cuda.current_stream(sim.device)
            for st in pp['streams']:
                cur.wait_stream(st)
            pp['primed'] = False
            self.update()

    def update(self):
        m = self.model
        if self.update_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self._update()
        if self.update_events is not None:
            e1.record()
            self.update_events.append((e0, e1))

    def _update(self):
        m = self.model
        boot = None
        if not self.done:
            _, boot, _ = m.forward(m.obs_slot(m.T), False, out_type='v')    # utils.py:190
        lr = self.lr.get(m.T) if hasattr(self.lr, "get") else self.lr
        beta = self.beta.get(m.T) if hasattr(self.beta, "get") else self.beta
        m.backward(boot, lr, beta)
        self.n_updates += 1
        if self.done:
            self.episode_rewards.append(float((self._rew_acc / self.T_episode).mean()))   # utils.py:296-305
            self.start_episode()

    def run(self, n_control_steps: int):
        for _ in range(n_control_steps):
            self.control_step()

"""IA2C / MA2C with the reference's constructor and method names (agents/models.py:132-261).

    IA2C(n_s_ls, n_a_ls, n_w_ls, total_step, model_config, seed=0)
    MA2C(n_s_ls, n_a_ls, n_w_ls, n_f_ls, total_step, model_config, seed=0)
    forward(obs, done, out_type='pv'), backward(R_ls, summary_writer=None, global_step=None),
    add_transition(obs, actions, rewards, values, done), reset(), save(dir, step), load(dir, checkpoint)
    attributes n_step, n_agent, sess (None: there is no TF session), policy_ls is not provided.

They are thin hosts over `BatchedA2C` (hand-written CUDA kernels); with `n_replicas == 1` the
methods take and return the reference's Python lists / numpy arrays, so `utils.py:Trainer` drives
them unchanged.  `batched` exposes the device-resident learner for R > 1.
"""
from __future__ import annotations

import logging
import os
from typing import List, Optional

import numpy as np
import torch

from .layout import PolicyLayout
from .learner import BatchedA2C
from .utils import Scheduler


class IA2C:
    name = 'ia2c'

    def __init__(self, n_s_ls, n_a_ls, n_w_ls, total_step, model_config, seed=0, n_f_ls=None,
                 n_replicas=1, device=0, obs_off=None, policy='lstm', **learner_kw):
        """policy='lstm': LstmACPolicy / FPLstmACPolicy, what the reference builds (agents/models.py:40-51);
        policy='fc': FcACPolicy (agents/policies.py:214-256), the FC variant of BASELINE config 2."""
        self.n_agent = len(n_s_ls)
        self.reward_clip = model_config.getfloat('reward_clip')
        self.reward_norm = model_config.getfloat('reward_norm')
        self.n_s_ls, self.n_a_ls, self.n_w_ls = list(n_s_ls), list(n_a_ls), list(n_w_ls)
        self.n_f_ls = list(n_f_ls) if n_f_ls is not None else [0] * self.n_agent
        self.n_step = model_config.getint('batch_size')
        self.sess = None
        self.total_step = total_step
        if obs_off is None:
            obs_off = np.concatenate([[0], np.cumsum(self.n_s_ls)])
        n_obs = int(obs_off[self.n_agent]) if len(obs_off) > self.n_agent else int(np.sum(self.n_s_ls))
        ff = model_config.getint('num_fp') if self.name == 'ma2c' else 0
        self.layout = PolicyLayout(self.n_s_ls, self.n_a_ls, self.n_w_ls, self.n_f_ls, obs_off, n_obs,
                                   fw=model_config.getint('num_fw'), ft=model_config.getint('num_ft'), ff=ff,
                                   h=model_config.getint('num_lstm'), recurrent=(policy != 'fc'))
        if policy == 'fc':
            from .learner_fc import BatchedFcA2C as _Learner
        else:
            _Learner = BatchedA2C
        self.batched = _Learner(
            self.layout, n_replicas, self.n_step, gamma=model_config.getfloat('gamma'),
            v_coef=model_config.getfloat('value_coef'), max_grad_norm=model_config.getfloat('max_grad_norm'),
            alpha=model_config.getfloat('rmsp_alpha'), eps=model_config.getfloat('rmsp_epsilon'),
            reward_norm=self.reward_norm, reward_clip=self.reward_clip, seed=seed, device=device, **learner_kw)
        if total_step:
            self._init_scheduler(model_config)
        self._rng = np.random.RandomState(seed)
        self._obs_dev = torch.zeros(n_replicas, n_obs, device=self.batched.dev)

    def _init_scheduler(self, model_config):                      # agents/models.py:53-69
        lr_init = model_config.getfloat('lr_init')
        lr_decay = model_config.get('lr_decay')
        beta_init = model_config.getfloat('entropy_coef_init')
        beta_decay = model_config.get('entropy_decay')
        if lr_decay == 'constant':
            self.lr_scheduler = Scheduler(lr_init, decay=lr_decay)
        else:
            self.lr_scheduler = Scheduler(lr_init, model_config.getfloat('LR_MIN'), self.total_step, decay=lr_decay)
        if beta_decay == 'constant':
            self.beta_scheduler = Scheduler(beta_init, decay=beta_decay)
        else:
            self.beta_scheduler = Scheduler(beta_init, model_config.getfloat('ENTROPY_COEF_MIN'),
                                            self.total_step * model_config.getfloat('ENTROPY_RATIO'),
                                            decay=beta_decay)

    # ---- reference protocol (lists in / lists out, one replica) ---------------------------------
    def _pack(self, obs: List[np.ndarray]) -> torch.Tensor:
        row = np.concatenate([np.asarray(o, np.float32) for o in obs])
        self._obs_dev[0, :row.shape[0]].copy_(torch.from_numpy(row))
        return self._obs_dev

    def forward(self, obs, done, out_type='pv'):
        b = self.batched
        slot = b.obs_slot() if ('p' in out_type and b.t < b.T) else self._obs_dev
        slot.copy_(self._pack(obs))
        pi, val, _ = b.forward(slot, bool(done), out_type, sample=False)
        pol = [pi[0, i, :self.n_a_ls[i]].cpu().numpy() for i in range(self.n_agent)] if 'p' in out_type else None
        vals = [float(v) for v in val[0].cpu().numpy()] if 'v' in out_type else None
        if len(out_type) == 1:
            return pol if out_type == 'p' else vals
        return pol, vals

    def add_transition(self, obs, actions, rewards, values, done):
        b = self.batched
        dev = b.dev
        act = torch.tensor(np.asarray(actions, np.int32).reshape(1, -1), device=dev)
        val = torch.tensor(np.asarray(values, np.float32).reshape(1, -1), device=dev)
        rew = torch.tensor(np.asarray(rewards, np.float32).reshape(1, -1) * np.ones((1, self.n_agent), np.float32),
                           device=dev)
        b.add_transition(rew, self._pre_done, bool(done), act=act, val=val)
        self._pre_done = bool(done)

    _pre_done = False

    def backward(self, R_ls, summary_writer=None, global_step=None):
        cur_lr = self.lr_scheduler.get(self.n_step)
        cur_beta = self.beta_scheduler.get(self.n_step)
        boot = torch.tensor(np.asarray(R_ls, np.float32).reshape(1, -1), device=self.batched.dev)
        self.batched.backward(boot, cur_lr, cur_beta)

    def reset(self):
        self.batched.reset()

    # ---- checkpoints: same file-name convention as the reference (agents/models.py:83-108) -------
    def save(self, model_dir, global_step):
        b = self.batched
        torch.save({'params': b.P.cpu(), 'rms': b.MS.cpu(), 'step': int(global_step), 'name': self.name},
                   os.path.join(model_dir, 'checkpoint-%d.pt' % int(global_step)))

    def load(self, model_dir, checkpoint=None):
        save_file, save_step = None, 0
        if os.path.exists(model_dir):
            if checkpoint is None:
                for file in os.listdir(model_dir):
                    if file.startswith('checkpoint'):
                        tokens = file.split('.')[0].split('-')
                        if len(tokens) != 2:
                            continue
                        if int(tokens[1]) > save_step:
                            save_file, save_step = file, int(tokens[1])
            else:
                save_file = 'checkpoint-%d.pt' % int(checkpoint)
        if save_file is not None and os.path.exists(os.path.join(model_dir, save_file)):
            ck = torch.load(os.path.join(model_dir, save_file))
            self.batched.P.copy_(ck['params']); self.batched.MS.copy_(ck['rms'])
            self.batched.pack_weights()
            logging.info('Checkpoint loaded: %s' % save_file)
            return True
        logging.error('Can not find old checkpoint for %s' % model_dir)
        return False


class MA2C(IA2C):
    name = 'ma2c'

    def __init__(self, n_s_ls, n_a_ls, n_w_ls, n_f_ls, total_step, model_config, seed=0, **kw):
        super().__init__(n_s_ls, n_a_ls, n_w_ls, total_step, model_config, seed=seed, n_f_ls=n_f_ls, **kw)


class IQL:
    """Independent Q-learning (agents/models.py:264-376, agents/policies.py:285-389): per-agent
    linear ('lr', LRQPolicy) or two-layer ('dqn', DeepQPolicy) Q network, epsilon-greedy exploration,
    1-step TD target WITHOUT a target network (policies.py:318-322), Adam, replay buffer, 10
    minibatches per agent per backward() (models.py:337-345).

    This is BASELINE config 1 ("reference plumbing", single env): it runs on PyTorch tensor ops
    (device tensors + torch.optim.Adam), not on hand-written kernels — it is not part of the measured
    hot path."""
    name = 'iql'

    def __init__(self, n_s_ls, n_a_ls, n_w_ls, total_step, model_config, seed=0, model_type='dqn', device=None):
        from .utils import ReplayBuffer
        self.model_type = model_type
        self.n_agent = len(n_s_ls)
        self.reward_clip = model_config.getfloat('reward_clip')
        self.reward_norm = model_config.getfloat('reward_norm')
        self.n_s_ls, self.n_a_ls, self.n_w_ls = list(n_s_ls), list(n_a_ls), list(n_w_ls)
        self.n_step = model_config.getint('batch_size')
        self.sess = None
        self.total_step = total_step
        self.dev = torch.device(device if device is not None else ('cuda' if torch.cuda.is_available() else 'cpu'))
        self.gamma = model_config.getfloat('gamma')
        self.max_grad_norm = model_config.getfloat('max_grad_norm')
        rng = np.random.RandomState(seed)
        self._np_rng = np.random.RandomState(seed + 1)
        from .layout import ortho_init
        self.nets, self.opts = [], []
        for n_s, n_a, n_w in zip(self.n_s_ls, self.n_a_ls, self.n_w_ls):
            layers = {}
            if model_type == 'dqn':
                n_fc, n_h = model_config.getint('num_fc'), model_config.getint('num_h')
                if n_w == 0:
                    layers['q_fcw'] = (n_s, n_fc); width = n_fc
                else:
                    layers['q_fcw'] = (n_s - n_w, n_fc); layers['q_fct'] = (n_w, n_fc // 4); width = n_fc + n_fc // 4
                layers['q_fc_0'] = (width, n_h); layers['q'] = (n_h, n_a)
            else:
                layers['q'] = (n_s, n_a)
            params = {}
            for k, shp in layers.items():
                params[k + '/w'] = torch.tensor(ortho_init(rng, shp), device=self.dev, requires_grad=True)
                params[k + '/b'] = torch.zeros(shp[1], device=self.dev, requires_grad=True)
            self.nets.append(params)
        if total_step:
            lr_init = model_config.getfloat('lr_init')
            lr_decay = model_config.get('lr_decay')
            self.lr_scheduler = Scheduler(lr_init, decay=lr_decay) if lr_decay == 'constant' else \
                Scheduler(lr_init, model_config.getfloat('lr_min'), total_step, decay=lr_decay)
            eps_init = model_config.getfloat('epsilon_init')
            eps_decay = model_config.get('epsilon_decay')
            self.eps_scheduler = Scheduler(eps_init, decay=eps_decay) if eps_decay == 'constant' else \
                Scheduler(eps_init, model_config.getfloat('epsilon_min'),
                          total_step * model_config.getfloat('epsilon_ratio'), decay=eps_decay)
            buffer_size = model_config.getfloat('buffer_size')
            self.trans_buffer_ls = [ReplayBuffer(buffer_size, self.n_step) for _ in range(self.n_agent)]
            self.opts = [torch.optim.Adam(list(p.values()), lr=lr_init) for p in self.nets]

    def _q(self, i, S):
        p, n_w = self.nets[i], self.n_w_ls[i]
        if self.model_type != 'dqn':
            return S @ p['q/w'] + p['q/b']
        if n_w == 0:
            h = torch.relu(S @ p['q_fcw/w'] + p['q_fcw/b'])
        else:
            n_s = S.shape[1] - n_w
            h = torch.cat([torch.relu(S[:, :n_s] @ p['q_fcw/w'] + p['q_fcw/b']),
                           torch.relu(S[:, n_s:] @ p['q_fct/w'] + p['q_fct/b'])], 1)
        h = torch.relu(h @ p['q_fc_0/w'] + p['q_fc_0/b'])
        return h @ p['q/w'] + p['q/b']

    def forward(self, obs, mode='act', stochastic=False):
        if mode == 'explore':
            eps = self.eps_scheduler.get(1)
        action, qs_ls = [], []
        for i in range(self.n_agent):
            with torch.no_grad():
                qs = self._q(i, torch.as_tensor(np.asarray(obs[i], np.float32)[None], device=self.dev))[0].cpu().numpy()
            if mode == 'explore' and self._np_rng.random_sample() < eps:
                action.append(int(self._np_rng.randint(self.n_a_ls[i])))
            elif not stochastic:
                action.append(int(np.argmax(qs)))
            else:
                pq = qs / np.sum(qs)
                action.append(int(self._np_rng.choice(np.arange(len(pq)), p=pq)))
            qs_ls.append(qs)
        return action, qs_ls

    def add_transition(self, obs, actions, rewards, next_obs, done):
        rewards = np.asarray(rewards, np.float64)
        if self.reward_norm:
            rewards = rewards / self.reward_norm
        if self.reward_clip:
            rewards = np.clip(rewards, -self.reward_clip, self.reward_clip)
        for i in range(self.n_agent):
            self.trans_buffer_ls[i].add_transition(obs[i], actions[i], rewards[i], next_obs[i], done)

    def backward(self, summary_writer=None, global_step=None):
        cur_lr = self.lr_scheduler.get(self.n_step)
        if self.trans_buffer_ls[0].size < self.trans_buffer_ls[0].batch_size:
            return
        for i in range(self.n_agent):
            for g in self.opts[i].param_groups:
                g['lr'] = cur_lr
            for _ in range(10):
                obs, acts, next_obs, rs, dones = self.trans_buffer_ls[i].sample_transition()
                S = torch.as_tensor(obs.astype(np.float32), device=self.dev)
                S1 = torch.as_tensor(next_obs.astype(np.float32), device=self.dev)
                A = torch.as_tensor(acts.astype(np.int64), device=self.dev)
                R = torch.as_tensor(rs.astype(np.float32), device=self.dev)
                D = torch.as_tensor(dones.astype(bool), device=self.dev)
                q0 = self._q(i, S).gather(1, A[:, None])[:, 0]
                with torch.no_grad():
                    tq = torch.where(D, R, R + self.gamma * self._q(i, S1).max(1)[0])
                loss = ((q0 - tq) ** 2).mean()
                self.opts[i].zero_grad()
                loss.backward()
                if self.max_grad_norm > 0:
                    torch.nn.utils.clip_grad_norm_(list(self.nets[i].values()), self.max_grad_norm)
                self.opts[i].step()

    def reset(self):
        return

    def save(self, model_dir, global_step):
        torch.save({'nets': [{k: v.detach().cpu() for k, v in p.items()} for p in self.nets], 'step': int(global_step)},
                   os.path.join(model_dir, 'checkpoint-%d.pt' % int(global_step)))

    def load(self, model_dir, checkpoint=None):
        files = [f for f in os.listdir(model_dir)] if os.path.exists(model_dir) else []
        steps = [int(f.split('.')[0].split('-')[1]) for f in files if f.startswith('checkpoint-')]
        if checkpoint is None and not steps:
            logging.error('Can not find old checkpoint for %s' % model_dir)
            return False
        step = int(checkpoint) if checkpoint is not None else max(steps)
        ck = torch.load(os.path.join(model_dir, 'checkpoint-%d.pt' % step))
        for p, q in zip(self.nets, ck['nets']):
            for k in p:
                p[k].data.copy_(q[k])
        return True

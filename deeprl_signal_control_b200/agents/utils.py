"""Host-side helpers that keep the reference's names and call signatures (reference agents/utils.py):

* `Scheduler`     — learning-rate / entropy-coefficient schedule (reference agents/utils.py:268-281): `get(n)` advances
                    the step counter by n and returns the value; 'linear' decays to `val_min` over `total_step` steps,
                    anything else is constant.
* `ReplayBuffer`  — IQL's experience ring (reference agents/utils.py:231-263): fixed capacity, oldest entry overwritten,
                    uniform sampling without replacement; `sample_transition()` returns (obs, actions, next_obs, rewards,
                    dones) in that order.
"""
from __future__ import annotations

import random

import numpy as np


class Scheduler:
    def __init__(self, val_init, val_min=0, total_step=0, decay='linear'):
        self._v0, self._floor = val_init, val_min
        self._horizon = float(total_step)
        self._linear = (decay == 'linear')
        self._steps_seen = 0
        # attribute names the reference exposes
        self.val, self.val_min, self.N, self.decay = val_init, val_min, self._horizon, decay

    @property
    def n(self):
        return self._steps_seen

    def get(self, n_step):
        self._steps_seen += n_step
        if not self._linear:
            return self._v0
        frac_left = 1 - self._steps_seen / self._horizon
        return self._floor if self._v0 * frac_left < self._floor else self._v0 * frac_left


class ReplayBuffer:
    def __init__(self, buffer_size, batch_size):
        self.buffer_size, self.batch_size = int(buffer_size), int(batch_size)
        self.reset()

    def reset(self):
        self._slots = [None] * self.buffer_size      # (ob, a, r, next_ob, done) per slot
        self.cum_size = 0

    @property
    def size(self):
        return self.cum_size if self.cum_size < self.buffer_size else self.buffer_size

    @property
    def buffer(self):
        return self._slots[:self.size]

    def add_transition(self, ob, a, r, next_ob, done):
        self._slots[self.cum_size % self.buffer_size] = (ob, a, r, next_ob, done)
        self.cum_size += 1

    def sample_transition(self):
        picks = random.sample(range(self.size), self.batch_size)      # same index draws as sampling the list itself
        field = lambda k: np.asarray([self._slots[i][k] for i in picks])
        return field(0), field(1), field(3), field(2), field(4)

"""Host-side helpers with the reference's names (agents/utils.py)."""


class Scheduler:
    """agents/utils.py:268-281."""

    def __init__(self, val_init, val_min=0, total_step=0, decay='linear'):
        self.val = val_init
        self.N = float(total_step)
        self.val_min = val_min
        self.decay = decay
        self.n = 0

    def get(self, n_step):
        self.n += n_step
        if self.decay == 'linear':
            return max(self.val_min, self.val * (1 - self.n / self.N))
        return self.val


import random as _random

import numpy as _np


class ReplayBuffer:
    """agents/utils.py:231-263 (ring buffer of (ob, a, r, next_ob, done), uniform sampling)."""

    def __init__(self, buffer_size, batch_size):
        self.buffer_size = buffer_size
        self.batch_size = batch_size
        self.cum_size = 0
        self.buffer = []

    def add_transition(self, ob, a, r, next_ob, done):
        experience = (ob, a, r, next_ob, done)
        if self.cum_size < self.buffer_size:
            self.buffer.append(experience)
        else:
            self.buffer[int(self.cum_size % self.buffer_size)] = experience
        self.cum_size += 1

    def reset(self):
        self.buffer = []
        self.cum_size = 0

    def sample_transition(self):
        minibatch = _random.sample(self.buffer, self.batch_size)
        cols = list(zip(*minibatch))
        return (_np.asarray(cols[0]), _np.asarray(cols[1]), _np.asarray(cols[3]), _np.asarray(cols[2]),
                _np.asarray(cols[4]))

    @property
    def size(self):
        return min(self.buffer_size, self.cum_size)

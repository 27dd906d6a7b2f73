"""large_grid scenario classes with the reference's names (envs/large_grid_env.py)."""
from __future__ import annotations

import numpy as np

STATE_NAMES = ['wave', 'wait']      # envs/large_grid_env.py:19
PHASE_NUM = 5                       # envs/large_grid_env.py:20


class LargeGridController:
    """Greedy policy of the reference (envs/large_grid_env.py:45-60): per node, pick the phase
    whose green lanes carry the largest summed wave.  Accepts [A][>=6] observations (list of
    arrays, as the reference) or a batched [R, A, >=6] array."""

    def __init__(self, node_names):
        self.name = 'greedy'
        self.node_names = node_names

    def forward(self, obs):
        ob = np.asarray(obs) if not isinstance(obs, list) else np.stack([np.asarray(o)[:6] for o in obs])
        flows = np.stack([ob[..., 0] + ob[..., 3], ob[..., 2] + ob[..., 5], ob[..., 1] + ob[..., 4],
                          ob[..., 1] + ob[..., 2], ob[..., 4] + ob[..., 5]], axis=-1)
        return np.argmax(flows, axis=-1)

    def greedy(self, ob, node_name):
        return int(self.forward([ob])[0])

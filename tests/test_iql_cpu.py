"""CPU: IQL plumbing (BASELINE config 1): epsilon schedule, replay buffer, TD update direction."""
import configparser

import numpy as np

INI = """
[MODEL_CONFIG]
gamma = 0.99
lr_init = 1e-2
lr_decay = constant
epsilon_init = 1.0
epsilon_min = 0.01
epsilon_decay = linear
epsilon_ratio = 0.5
max_grad_norm = 40
batch_size = 20
buffer_size = 1000
reward_norm = 1.0
reward_clip = 2.0
num_fc = 16
num_h = 8
"""


def test_iql_lr_learns_a_bandit_and_follows_the_reference_protocol():
    from deeprl_signal_control_b200.agents.models import IQL
    cp = configparser.ConfigParser(); cp.read_string(INI)
    n_s_ls, n_a_ls, n_w_ls = [4, 6], [3, 2], [0, 0]
    m = IQL(n_s_ls, n_a_ls, n_w_ls, 400, cp["MODEL_CONFIG"], seed=0, model_type="lr", device="cpu")
    rng = np.random.default_rng(0)
    best = [2, 0]
    for step in range(400):
        obs = [rng.random(n).astype(np.float32) for n in n_s_ls]
        act, qs = m.forward(obs, mode="explore")
        assert len(act) == 2 and qs[0].shape == (3,)
        rew = np.array([1.0 if act[i] == best[i] else -1.0 for i in range(2)])
        m.add_transition(obs, act, rew, obs, True)          # done -> target = r (bandit)
        if step % 20 == 19:
            m.backward(None, step)
    # epsilon decays linearly over total_step * ratio (agents/models.py:306-316)
    assert abs(m.eps_scheduler.get(0) - 0.01) < 1e-9
    obs = [rng.random(n).astype(np.float32) for n in n_s_ls]
    act, _ = m.forward(obs)                                  # greedy
    assert act == best
    m2 = IQL(n_s_ls, n_a_ls, [0, 2], 100, cp["MODEL_CONFIG"], seed=0, model_type="dqn", device="cpu")
    a2, q2 = m2.forward(obs)
    assert q2[1].shape == (2,)

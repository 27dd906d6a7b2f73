"""GPU: episode boundaries in the device-resident loop (utils.py:277-305): done flags, env/model reset,
bootstrap skipped at episode end, per-episode reward log."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_short_episodes_reset_and_log():
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    from deeprl_signal_control_b200.agents.learner import BatchedA2C
    from deeprl_signal_control_b200.agents.trainer import BatchedTrainer
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from deeprl_signal_control_b200.sim import BatchedSim
    net = build_large_grid(agent="ma2c", episode_length_sec=300)
    par = EnvParams(agent="ma2c", episode_length_sec=300)            # 60 control steps per episode
    R = 32
    sim = BatchedSim(net, par, R)
    lay = PolicyLayout(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, net.node_obs_off, net.n_obs, fw=128, ft=32, ff=64)
    model = BatchedA2C(lay, R, n_step=30, reward_norm=2000.0, reward_clip=2.0, seed=3, chunk=32)
    tr = BatchedTrainer(sim, model, "ma2c", lr=5e-4, beta=0.01, seed0=12)
    assert tr.T_episode == 60
    seen_done_pre = []
    for step in range(150):
        seen_done_pre.append(tr.done)
        tr.control_step()
    torch.cuda.synchronize()
    # pre-decision done is True exactly on the first step of every episode (utils.py:279-281)
    assert [i for i, d in enumerate(seen_done_pre) if d] == [0, 60, 120]
    assert tr.n_updates == 5 and len(tr.episode_rewards) == 2
    assert all(np.isfinite(tr.episode_rewards)) and all(r <= 0 for r in tr.episode_rewards)
    assert tr.step_in_episode == 30
    # a fresh episode starts from an empty network: few live vehicles 30 steps in, yet some
    assert 1 < sim.mean_live() < 200
    assert torch.isfinite(model.P).all()

"""GPU: the device-resident MA2C / IA2C training loop end to end on the real 5x5 grid."""
import configparser

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MODEL_INI = """
[MODEL_CONFIG]
rmsp_alpha = 0.99
rmsp_epsilon = 1e-5
max_grad_norm = 40
gamma = 0.99
lr_init = 5e-4
lr_decay = constant
entropy_coef_init = 0.01
entropy_decay = constant
value_coef = 0.5
num_fw = 128
num_ft = 32
num_lstm = 64
num_fp = 64
batch_size = 120
reward_norm = 2000.0
reward_clip = 2.0
"""


@pytest.mark.parametrize("agent", ["ma2c", "ia2c"])
def test_training_loop_runs_and_updates(agent):
    from deeprl_signal_control_b200.agents.models import IA2C, MA2C
    from deeprl_signal_control_b200.agents.trainer import BatchedTrainer
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from deeprl_signal_control_b200.sim import BatchedSim
    cp = configparser.ConfigParser(); cp.read_string(MODEL_INI)
    net, par = build_large_grid(agent=agent), EnvParams(agent=agent)
    R = 64
    sim = BatchedSim(net, par, R)
    if agent == "ma2c":
        model = MA2C(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, 1e6, cp["MODEL_CONFIG"], seed=1,
                     n_replicas=R, obs_off=net.node_obs_off, chunk=48)
    else:
        model = IA2C(net.n_s_ls, net.n_a_ls, net.n_w_ls, 1e6, cp["MODEL_CONFIG"], seed=1,
                     n_replicas=R, obs_off=net.node_obs_off, chunk=48)
    b = model.batched
    tr = BatchedTrainer(sim, b, agent, lr=5e-4, beta=0.01)
    P0 = b.P.clone()
    tr.run(240)                     # two rollouts + two updates
    torch.cuda.synchronize()
    assert tr.n_updates == 2 and b.t == 0
    assert torch.isfinite(b.P).all() and not torch.equal(P0, b.P)
    assert float(b.norms.min()) > 0
    # fingerprints fed to the simulator are the policy probabilities: rows sum to 1
    np.testing.assert_allclose(b.pi.sum(-1).cpu().numpy(), 1.0, rtol=1e-5)
    # observation slot consumed by forward == observation produced by the simulator
    assert float(b.obs_hist[0].abs().sum()) > 0
    assert sim.mean_live() > 20


def test_pipelined_host_step_matches_device_loop():
    """control_step_host_pipelined (replica ranges on separate streams through tsc_step_host_range /
    tscl_policy_step_v2r) reproduces the device-resident loop bit for bit up to the first update, and keeps running
    across the update / re-priming boundary."""
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    from deeprl_signal_control_b200.agents.learner import BatchedA2C
    from deeprl_signal_control_b200.agents.trainer import BatchedTrainer
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from deeprl_signal_control_b200.sim import BatchedSim
    agent, R, T = "ma2c", 300, 12
    net, par = build_large_grid(agent=agent), EnvParams(agent=agent)
    lay = PolicyLayout(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, net.node_obs_off, net.n_obs, fw=128, ft=32,
                       ff=64, h=64)
    trs = []
    for _ in range(3):
        sim = BatchedSim(net, par, R)
        m = BatchedA2C(lay, R, n_step=T, reward_norm=2000.0, reward_clip=2.0, seed=3, chunk=150)
        trs.append(BatchedTrainer(sim, m, agent, lr=5e-4, beta=0.01, seed0=7))
    dev, host, pipe = trs
    for _ in range(T - 1):
        dev.control_step(); host.control_step_host(); pipe.control_step_host_pipelined(n_parts=3)
    torch.cuda.synchronize()
    for other in (host, pipe):
        assert torch.equal(dev.model.act_hist[:T - 1], other.model.act_hist[:T - 1])
        assert torch.equal(dev.model.rew_hist[:T - 1], other.model.rew_hist[:T - 1])
        assert torch.equal(dev.model.obs_hist[:T], other.model.obs_hist[:T])
        assert torch.equal(dev.model.val_hist[:T - 1], other.model.val_hist[:T - 1])
    assert torch.equal(dev._rew_acc, pipe._rew_acc)      # (the pipelined loop's recurrent state is one decision ahead)
    assert all(pipe.model._acts_ok[:T - 1]) == all(dev.model._acts_ok[:T - 1])
    # across the update and the re-priming after it
    for _ in range(T + 3):
        dev.control_step(); pipe.control_step_host_pipelined(n_parts=3)
    torch.cuda.synchronize()
    assert dev.n_updates == pipe.n_updates == 2 and dev.model.t == pipe.model.t
    assert pipe.model.n_forward == dev.model.n_forward + 1      # the pipelined loop has already issued the next decision
    assert torch.isfinite(pipe.model.P).all()
    # gradients are accumulated with fp32 atomics (order-dependent), so parameters agree to rounding only
    assert float((dev.model.P - pipe.model.P).abs().max()) < 1e-4

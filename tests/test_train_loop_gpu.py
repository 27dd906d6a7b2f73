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

"""CPU tests of the learner's host logic and of its oracle restatement against the reference's own
pure-Python pieces (tests/golden/buffers.npz from agents/utils.py)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_returns_restatement_matches_reference_buffer():
    from oracle.learner_ref import nstep_returns
    z = np.load(os.path.join(GOLD, "buffers.npz"))
    for b in range(z["buf_rs"].shape[0]):
        Rs, Advs = nstep_returns(list(z["buf_rs"][b]), list(z["buf_vs"][b]), list(z["buf_dones"][b].astype(float)),
                                 float(z["buf_R"][b]), 0.99)
        np.testing.assert_allclose(Rs, z["buf_Rs"][b], rtol=1e-6, atol=1e-6)     # reference casts to f32
        np.testing.assert_allclose(Advs, z["buf_Advs"][b], rtol=1e-6, atol=1e-6)
    # pre-step dones of batch b start with the last post-step done of batch b-1 (agents/utils.py:187-193,226)
    assert bool(z["buf_pre_dones"][0][0]) is True
    for b in range(1, z["buf_rs"].shape[0]):
        assert bool(z["buf_pre_dones"][b][0]) == bool(z["buf_dones"][b - 1][-1])
        assert list(z["buf_pre_dones"][b][1:]) == list(z["buf_dones"][b][:-1])


def test_scheduler_matches_reference():
    from deeprl_signal_control_b200.agents.utils import Scheduler
    z = np.load(os.path.join(GOLD, "buffers.npz"))
    s = Scheduler(1.0, 0.01, 1000.0, decay="linear")
    np.testing.assert_allclose([s.get(120) for _ in range(12)], z["sched_linear"], rtol=1e-12)
    s = Scheduler(5e-4, decay="constant")
    np.testing.assert_allclose([s.get(120) for _ in range(3)], z["sched_const"], rtol=1e-12)


def test_layout_matches_reference_parameter_count():
    """MA2C grid: 3.93 M parameters, IA2C grid: 3.07 M (SURVEY §2a / BASELINE.md §3)."""
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    for agent, ff, want in (("ma2c", 64, 3931990), ("ia2c", 0, 3068630)):
        net = build_large_grid(agent=agent)
        lay = PolicyLayout(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, net.node_obs_off, net.n_obs,
                           fw=128, ft=32, ff=ff, h=64)
        # our heads are padded to max_na: V heads carry max_na-1 unused (always-zero) columns per row
        pad = lay.A * (lay.h + 1) * (lay.max_na - 1)
        assert lay.n_params - pad == want
        assert lay.agent_of.max() == lay.A - 1 and len(lay.agent_of) == lay.n_params
        P = lay.init_params(0)
        v = lay.views(P)
        w = v["wx"][3].astype(np.float64)
        np.testing.assert_allclose(w.T @ w if w.shape[0] >= w.shape[1] else w @ w.T,
                                   2 * np.eye(min(w.shape)), atol=1e-4)   # orthogonal, scale sqrt(2)
        assert np.all(v["bl"] == 0)


def test_fc_policy_layout_matches_reference_variable_shapes():
    """FcACPolicy (agents/policies.py:214-256): per net fcw [n_wave,128]+b, fct [n_wait,32]+b, fc [160,64]+b, head."""
    import torch
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    from oracle.learner_ref import unit_forward
    n_w, n_wave, n_a = [6, 6], [18, 30], [5, 4]
    n_s = [w + t for w, t in zip(n_wave, n_w)]
    off = np.concatenate([[0], np.cumsum(n_s)]).astype(np.int32)
    lay = PolicyLayout(n_s, n_a, n_w, [0, 0], off, int(off[-1]), fw=128, ft=32, ff=0, h=64, recurrent=False)
    assert lay.dx == 160 and not lay.recurrent
    v = lay.views(lay.init_params(3))
    assert v["wx"].shape == (4, 160, 64) and v["wh"].size == 0 and v["bl"].shape == (4, 64)
    expect = sum(2 * (nw * 128 + 128 + nt * 32 + 32 + 160 * 64 + 64 + 64 * 5 + 5) for nw, nt in zip(n_wave, n_w))
    assert lay.n_params == expect
    # float64 restatement runs and normalises
    P = torch.from_numpy(lay.init_params(3).astype(np.float64))
    obs = torch.rand(2, 7, lay.n_obs, dtype=torch.float64)
    pi = unit_forward(lay.views(P), lay, 0, obs, [0.0, 0.0], None, None)[0]
    val = unit_forward(lay.views(P), lay, 1, obs, [0.0, 0.0], None, None)[0]
    assert pi.shape == (2, 7, 5) and val.shape == (2, 7)
    np.testing.assert_allclose(pi.sum(-1).numpy(), 1.0, rtol=1e-12)

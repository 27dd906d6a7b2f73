"""CPU: the learner oracle (oracle/learner_ref.py), the checkpoint name map and the IQL host code against vectors
produced by the REFERENCE's own agents/{utils,policies,models}.py executed on the TF1 shim
(tests/golden/gen_learner_golden.py -> tests/golden/learner_*.npz).  This is what pins rows a11 / a12 / a15 / a16:
the oracle is no longer an unpinned reading of the reference, it reproduces what the reference's graph builders
compute (float64; weights / gradients are stored as float32 in the fixtures, hence the 1e-6-level tolerances)."""
import configparser
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _layout(z, recurrent=True):
    from deeprl_signal_control_b200.agents.layout import PolicyLayout
    n_wave, n_wait, n_a = z["n_wave"], z["n_wait"], z["n_a"]
    n_fp = z["n_fp"] if "n_fp" in z.files else np.zeros_like(n_wave)
    n_s = [int(w + t + f) for w, t, f in zip(n_wave, n_wait, n_fp)]
    off = np.concatenate([[0], np.cumsum(n_s)]).astype(np.int32)
    ff = 64 if int(n_fp.sum()) > 0 else 0
    return PolicyLayout(n_s, [int(a) for a in n_a], [int(t) for t in n_wait], [int(f) for f in n_fp], off,
                        int(off[-1]), fw=128, ft=32, ff=ff, h=64, max_na=int(max(n_a)), recurrent=recurrent)


def _named(z, prefix):
    return {k[len(prefix) + 1:]: z[k] for k in z.files if k.startswith(prefix + "/")}


@pytest.mark.parametrize("kind,recurrent", [("ma2c", True), ("ia2c", True), ("fc", False)])
def test_variable_names_and_shapes_equal_the_reference_graph(kind, recurrent):
    from deeprl_signal_control_b200.agents import checkpoint as ck
    z = np.load(os.path.join(GOLD, "learner_%s.npz" % kind))
    lay = _layout(z, recurrent)
    specs = ck.variable_specs(lay)
    assert [s[0] for s in specs] == [str(n) for n in z["var_names"]]
    w0 = _named(z, "w0")
    for name, _, _, shape in specs:
        assert tuple(w0[name].shape) == tuple(shape), name
    # round trip through the flat layout
    flat = ck.import_named(lay, w0)
    back = ck.export_named(lay, flat)
    for name in w0:
        assert np.array_equal(back[name], w0[name]), name
    with pytest.raises(KeyError):
        ck.import_named(lay, {k: v for k, v in w0.items() if not k.endswith("pi/w")})
    bad = dict(w0); first = specs[0][0]; bad[first] = bad[first][:-1]
    with pytest.raises(ValueError):
        ck.import_named(lay, bad)


def _grad_close(lay, G, named_g, tol=2e-6):
    from deeprl_signal_control_b200.agents import checkpoint as ck
    got = ck.export_named(lay, G)
    for name, ref in named_g.items():
        scale = max(float(np.abs(ref).max()), 1e-8)
        err = float(np.abs(got[name].astype(np.float64) - ref).max()) / scale
        assert err < tol, (name, err)


@pytest.mark.parametrize("kind", ["ma2c", "ia2c"])
def test_oracle_reproduces_reference_lstm_policies_loss_gradients_and_updates(kind):
    """Two n-step updates following Trainer.explore (utils.py:142-193): per-step pi / v / LSTM states, n-step returns,
    per-agent loss, gradients, global norms, and the weights after RMSProp — all equal to what the reference's
    LstmACPolicy / FPLstmACPolicy graphs produced."""
    from deeprl_signal_control_b200.agents import checkpoint as ck
    from oracle.learner_ref import a2c_loss, clip_rmsprop, nstep_returns, unit_forward
    z = np.load(os.path.join(GOLD, "learner_%s.npz" % kind))
    lay = _layout(z)
    T, A, U = int(z["n_step"]), lay.A, lay.U
    gamma, v_coef, beta, lr = float(z["gamma"]), float(z["v_coef"]), float(z["beta"]), float(z["lr"])
    P = ck.import_named(lay, _named(z, "w0"), dtype=np.float64)
    MS = np.ones_like(P)
    variants = [("", float(z["max_grad_norm"]))]
    if "clip/max_grad_norm" in z.files:
        variants.append(("clip/", float(z["clip/max_grad_norm"])))
    c = [torch.zeros(1, 64, dtype=torch.float64) for _ in range(U)]
    h = [torch.zeros(1, 64, dtype=torch.float64) for _ in range(U)]
    for b in range(2):
        obs = torch.from_numpy(z["b%d/obs" % b])[:, None, :]              # [T, 1, n_obs]
        dpre, dpost = list(z["b%d/done_pre" % b]), list(z["b%d/done_post" % b])
        c_bw, h_bw = [x.clone() for x in c], [x.clone() for x in h]         # states_bw (agents/policies.py:153)
        v = lay.views(torch.from_numpy(P))
        for t in range(T):
            for a in range(A):
                na = int(lay.n_a[a])
                pi, _, c[2 * a], h[2 * a] = unit_forward(v, lay, 2 * a, obs[t:t + 1], [dpre[t]], c[2 * a], h[2 * a])
                val, _, c[2 * a + 1], h[2 * a + 1] = unit_forward(v, lay, 2 * a + 1, obs[t:t + 1], [dpre[t]],
                                                                  c[2 * a + 1], h[2 * a + 1])
                np.testing.assert_allclose(pi[0, 0].numpy(), z["b%d/pi" % b][t, a, :na], rtol=1e-7, atol=1e-10)
                np.testing.assert_allclose(val[0, 0].numpy(), z["b%d/val" % b][t, a], rtol=1e-7, atol=1e-10)
                st = z["b%d/states" % b][t, a]                              # [pi|v][c(64) | h(64)]
                np.testing.assert_allclose(c[2 * a][0].numpy(), st[0, :64], rtol=1e-7, atol=1e-10)
                np.testing.assert_allclose(h[2 * a + 1][0].numpy(), st[1, 64:], rtol=1e-7, atol=1e-10)
        # reward normalisation / clip of add_transition (agents/models.py:222-229) + OnPolicyBuffer returns
        rew = np.clip(z["b%d/rew" % b] / float(z["reward_norm"]), -float(z["reward_clip"]), float(z["reward_clip"]))
        Rs, Advs = nstep_returns(list(rew), list(z["b%d/val" % b]), dpost, z["b%d/boot" % b], gamma)
        np.testing.assert_allclose(Rs, z["b%d/Rs" % b], rtol=2e-6, atol=2e-6)          # reference casts to float32
        np.testing.assert_allclose(Advs, z["b%d/Advs" % b], rtol=2e-6, atol=2e-6)
        # pre-step dones feed the BPTT (agents/utils.py:226).  The buffer's very first entry is its constructor default
        # False while Trainer.run passes done=True to the first forward (utils.py:279); the two agree numerically because
        # model.reset() has zeroed the LSTM state there.
        dbw = list(z["b%d/dones_bw_0" % b])
        assert dbw[1:] == dpre[1:]
        if dbw[0] != dpre[0]:
            assert all(float(x.abs().max()) == 0.0 for x in c_bw + h_bw)
        Pt = torch.from_numpy(P).requires_grad_(True)
        loss, parts = a2c_loss(Pt, lay, obs, torch.from_numpy(z["b%d/acts" % b])[:, None, :],
                               torch.from_numpy(z["b%d/Rs" % b].astype(np.float64))[:, None, :],
                               torch.from_numpy(z["b%d/Advs" % b].astype(np.float64))[:, None, :], dbw, c_bw, h_bw,
                               v_coef, beta)
        np.testing.assert_allclose([sum(p) for p in parts], z["b%d/loss" % b], rtol=1e-7)
        loss.backward()
        G = Pt.grad.numpy()
        if b == 0:
            _grad_close(lay, G, _named(z, "b0/g"))
            for pre, mgn in variants[1:]:
                P1c, _, nrm = clip_rmsprop(P, G, MS, lay.agent_of, mgn, lr, float(z["alpha"]), float(z["eps"]), A)
                assert (nrm > mgn).all()                                                # the clip branch is active
                got = ck.export_named(lay, P1c)
                for name, ref in _named(z, "clip/w1").items():
                    np.testing.assert_allclose(got[name], ref, rtol=0, atol=1.5e-7, err_msg=name)
        P, MS, norms = clip_rmsprop(P, G, MS, lay.agent_of, variants[0][1], lr, float(z["alpha"]), float(z["eps"]), A)
        np.testing.assert_allclose(norms, z["b%d/grad_norm" % b], rtol=1e-7)
    got = ck.export_named(lay, P)
    for name, ref in _named(z, "w2").items():
        np.testing.assert_allclose(got[name], ref, rtol=0, atol=1.5e-7, err_msg=name)
        assert float(np.abs(ref - z["w0/" + name]).max()) > 2e-5 or name.endswith("/b") or ref.size < 8


def test_oracle_reproduces_reference_fc_policy():
    """FcACPolicy (agents/policies.py:214-256) driven directly: pi / v, loss, gradients, two RMSProp steps."""
    from deeprl_signal_control_b200.agents import checkpoint as ck
    from oracle.learner_ref import a2c_loss, clip_rmsprop, unit_forward
    z = np.load(os.path.join(GOLD, "learner_fc.npz"))
    lay = _layout(z, recurrent=False)
    T, A = int(z["n_step"]), lay.A
    P = ck.import_named(lay, _named(z, "w0"), dtype=np.float64)
    MS = np.ones_like(P)
    for b in range(2):
        obs = torch.from_numpy(z["b%d/obs" % b])[:, None, :]
        v = lay.views(torch.from_numpy(P))
        for a in range(A):
            na = int(lay.n_a[a])
            pi = unit_forward(v, lay, 2 * a, obs, [0.0] * T, None, None)[0]
            val = unit_forward(v, lay, 2 * a + 1, obs, [0.0] * T, None, None)[0]
            np.testing.assert_allclose(pi[:, 0].numpy(), z["b%d/pi" % b][:, a, :na], rtol=1e-7, atol=1e-10)
            np.testing.assert_allclose(val[:, 0].numpy(), z["b%d/val" % b][:, a], rtol=1e-7, atol=1e-10)
        Pt = torch.from_numpy(P).requires_grad_(True)
        loss, parts = a2c_loss(Pt, lay, obs, torch.from_numpy(z["b%d/acts" % b])[:, None, :],
                               torch.from_numpy(z["b%d/Rs" % b].astype(np.float64))[:, None, :],
                               torch.from_numpy(z["b%d/Advs" % b].astype(np.float64))[:, None, :], [0.0] * T,
                               [None] * lay.U, [None] * lay.U, float(z["v_coef"]), float(z["beta"]))
        np.testing.assert_allclose([sum(p) for p in parts], z["b%d/loss" % b], rtol=1e-7)
        loss.backward()
        G = Pt.grad.numpy()
        _grad_close(lay, G, _named(z, "b%d/g" % b))
        P, MS, norms = clip_rmsprop(P, G, MS, lay.agent_of, float(z["max_grad_norm"]), float(z["lr"]), float(z["alpha"]),
                                    float(z["eps"]), A)
        np.testing.assert_allclose(norms, z["b%d/grad_norm" % b], rtol=1e-7)
        got = ck.export_named(lay, P)
        for name, ref in _named(z, "w%d" % (b + 1)).items():
            np.testing.assert_allclose(got[name], ref, rtol=0, atol=1.5e-7, err_msg=name)


INI = """
[MODEL_CONFIG]
gamma = 0.99
lr_init = 1e-4
lr_decay = constant
epsilon_init = 1.0
epsilon_min = 0.01
epsilon_decay = linear
epsilon_ratio = 0.5
max_grad_norm = 40
batch_size = 20
buffer_size = 1000
reward_norm = 3.0
reward_clip = 2.0
num_fc = 128
num_h = 64
"""


@pytest.mark.parametrize("kind", ["lr", "dqn"])
def test_iql_td_update_matches_reference_qpolicy(kind):
    """LRQPolicy / DeepQPolicy (agents/policies.py:285-389): q-values, TD loss without a target network, global-norm
    clip and TF1 Adam over three consecutive minibatches; plus the variable names and the epsilon schedule."""
    from deeprl_signal_control_b200.agents.models import IQL
    z = np.load(os.path.join(GOLD, "learner_iql.npz"))
    cp = configparser.ConfigParser(); cp.read_string(INI)
    n_s_ls, n_a_ls = [24, 36], [5, 4]
    m = IQL(n_s_ls, n_a_ls, [0, 0], 1000, cp["MODEL_CONFIG"], seed=0, model_type=kind, device="cpu")
    w0 = _named(z, "%s/w0" % kind)
    assert sorted(m.named_weights().keys()) == sorted(str(n) for n in z["%s/var_names" % kind])
    m.load_named(w0)
    for k in range(3):
        for i in range(2):
            pre = "%s/k%d/a%d" % (kind, k, i)
            q = m._q(i, torch.from_numpy(z[pre + "/obs"].astype(np.float32))).detach().numpy()
            np.testing.assert_allclose(q, z[pre + "/q"], rtol=2e-5, atol=2e-6)
            loss, norm = m.td_update(i, z[pre + "/obs"], z[pre + "/acts"], z[pre + "/next_obs"], z[pre + "/dones"],
                                     z[pre + "/rs"], 1e-4)
            np.testing.assert_allclose(loss, float(z[pre + "/loss"]), rtol=2e-5)
            np.testing.assert_allclose(norm, float(z[pre + "/grad_norm"]), rtol=2e-5)
        got = m.named_weights()
        for name, ref in _named(z, "%s/w%d" % (kind, k + 1)).items():
            np.testing.assert_allclose(got[name], ref, rtol=0, atol=2e-6, err_msg=name)
            assert float(np.abs(ref - w0[name]).max()) > 5e-5
    if kind == "lr":
        ob = z["lr/plumb_obs"]
        acts, qs = m.forward([ob[:24], ob[24:]], mode="act")
        # the golden was taken on the trained net; ours has had the same three updates
        np.testing.assert_allclose(np.concatenate(qs), z["lr/plumb_q"], rtol=1e-4, atol=1e-5)
        assert acts == [int(a) for a in z["lr/plumb_act"]]
        eps = []
        for _ in range(5):
            m.forward([ob[:24], ob[24:]], mode="explore")
            eps.append(m.eps_scheduler.val * (1 - m.eps_scheduler.n / m.eps_scheduler.N))
        np.testing.assert_allclose(eps, z["lr/eps"], rtol=1e-12)


def test_checkpoint_files_follow_the_reference_naming(tmp_path):
    """save()/load() of the IQL host (CPU-capable) write `checkpoint-<step>.npz` keyed by the TF variable names and pick
    the highest step, as agents/models.py:83-108 does."""
    from deeprl_signal_control_b200.agents.models import IQL
    cp = configparser.ConfigParser(); cp.read_string(INI)
    m = IQL([24, 36], [5, 4], [0, 0], 1000, cp["MODEL_CONFIG"], seed=1, model_type="dqn", device="cpu")
    d = str(tmp_path) + "/"
    m.save(d, 100)
    w100 = {k: v.copy() for k, v in m.named_weights().items()}
    for p in m.nets:
        for v in p.values():
            v.data.add_(1.0)
    m.save(d, 2000)
    assert sorted(os.listdir(d)) == ["checkpoint-100.npz", "checkpoint-2000.npz"]
    z = np.load(d + "checkpoint-100.npz")
    assert "dqn_0a_q/q_fcw/w" in z.files and "dqn_1a_q/q/b" in z.files
    m2 = IQL([24, 36], [5, 4], [0, 0], 0, cp["MODEL_CONFIG"], seed=5, model_type="dqn", device="cpu")
    assert m2.load(d) is True                                     # highest step
    np.testing.assert_allclose(m2.named_weights()["dqn_0a_q/q/w"], w100["dqn_0a_q/q/w"] + 1.0)
    assert m2.load(d, checkpoint=100) is True
    np.testing.assert_array_equal(m2.named_weights()["dqn_0a_q/q/w"], w100["dqn_0a_q/q/w"])
    assert m2.load(str(tmp_path) + "/nope/") is False

"""GPU: Monaco scenario through the C ABI — bit-exact vs the oracle, and the reference-facing RealNetEnv."""
import configparser
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

REAL_INI = """
[ENV_CONFIG]
clip_wave = 2.0
clip_wait = 2.0
control_interval_sec = 5
agent = %s
coop_gamma = 0.9
data_path = ./real_net/data/
episode_length_sec = 3600
norm_wave = 5.0
norm_wait = 100.0
coef_wait = 0
flow_rate = 325
objective = queue
scenario = real_net
seed = 42
test_seeds = 10000,20000,30000
yellow_interval_sec = 2
"""


def test_real_net_bit_exact_vs_oracle():
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    from deeprl_signal_control_b200.sim import BatchedSim
    from oracle.sim_ref import RefSim
    from tests.test_real_net_cpu import real_params
    net, par = real_net_tables("ma2c"), real_params("ma2c")
    R = 5
    gpu, ref = BatchedSim(net, par, R), RefSim(net, par, R)
    seeds = np.arange(R, dtype=np.uint64) * np.uint64(101) + np.uint64(42)
    gpu.reset(seeds); ref.reset(seeds)
    rng = np.random.default_rng(5)
    na = np.array(net.n_a_ls)
    for step in range(400):
        act = (rng.integers(0, 1 << 20, (R, net.n_nodes)) % na).astype(np.int32)
        fp = rng.random((R, net.n_nodes, net.max_na), dtype=np.float32)
        obs, rew, grew, done = gpu.step(torch.from_numpy(act).cuda(), torch.from_numpy(fp).cuda())
        o2, r2, g2, d2 = ref.step(act, fp)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), o2.view(np.uint32))
        np.testing.assert_array_equal(rew.cpu().numpy().view(np.uint32), r2.view(np.uint32))
        np.testing.assert_array_equal(grew.cpu().numpy().view(np.uint32), g2.view(np.uint32))
        for a, b in zip(gpu.counts(), ref.counts()):
            np.testing.assert_array_equal(a.cpu().numpy(), b)
        if step % 50 == 0:
            for r in (0, R - 1):
                c1, v1 = gpu.dump_state(r); c2, v2 = ref.dump_state(r)
                np.testing.assert_array_equal(c1, c2); np.testing.assert_array_equal(v1, v2)
    assert ref.misc(0)["live"] > 30 and ref.misc(0)["arrived"] > 100


def test_real_net_env_reproduces_reference_trace():
    from deeprl_signal_control_b200.envs.real_net_env import RealNetEnv
    z = np.load(os.path.join(GOLD, "real_ma2c_train.npz"))
    meta = json.loads(str(z["meta"]))
    cp = configparser.ConfigParser()
    cp.read_string(REAL_INI % "ma2c")
    env = RealNetEnv(cp["ENV_CONFIG"])
    assert env.n_s_ls == meta["n_s_ls"] and env.n_a_ls == meta["n_a_ls"] and env.n_f_ls == meta["n_f_ls"]
    ob = env.reset()
    np.testing.assert_allclose(np.concatenate(ob), z["obs"][0], rtol=2e-6, atol=1e-6)
    for t in range(60):
        env.update_fingerprint([z["fps"][t][i, :env.n_a_ls[i]] for i in range(28)])
        ob, reward, done, greward = env.step(list(z["actions"][t]))
        np.testing.assert_allclose(np.concatenate(ob), z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(reward, z["reward"][t], rtol=3e-6, atol=1e-5)
        assert abs(greward - z["greward"][t]) < 1e-4

"""GPU: 6-intersection `small_grid` scenario (stochastic demand: JTRRouter turn ratios + `probability=` flows restated
as counter-RNG draws, include/tsc.h) — CUDA kernel vs the CPU oracle bit for bit, and `SmallGridEnv` replaying the
trace recorded from the reference's own SmallGridEnv (tests/golden/gen_small_grid_golden.py)."""
import configparser
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

SMALL_INI = """
[ENV_CONFIG]
clip_wave = 1000.0
clip_wait = 1000.0
control_interval_sec = 5
agent = %s
coop_gamma = 0.75
data_path = ./small_grid/data/
episode_length_sec = 3600
norm_wave = 1.0
norm_wait = 1.0
coef_wait = 0.2
num_extra_car_per_hour = 1000
objective = hybrid
scenario = small_grid
seed = 42
test_seeds = 10000,20000,30000
yellow_interval_sec = 2
"""


@pytest.mark.parametrize("agent", ["greedy", "ma2c"])
def test_small_grid_bit_exact_vs_oracle(agent):
    from deeprl_signal_control_b200.net.small_grid import build_small_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from deeprl_signal_control_b200.sim import BatchedSim
    from oracle.sim_ref import RefSim
    net = build_small_grid(agent=agent)
    par = EnvParams(agent=agent, norm_wave=1.0, norm_wait=1.0, clip_wave=1000.0, clip_wait=1000.0, coop_gamma=0.75)
    R = 5
    gpu, ref = BatchedSim(net, par, R), RefSim(net, par, R)
    seeds = np.arange(R, dtype=np.uint64) * np.uint64(104729) + np.uint64(3)
    gpu.reset(seeds); ref.reset(seeds)
    gpu.set_record(True); ref.set_record(True)
    rng = np.random.default_rng(4)
    n_a = np.asarray(net.n_a_ls)
    for step in range(720):                                  # one whole episode, all six demand intervals
        act = (rng.integers(0, 1 << 30, size=(R, net.n_nodes)) % n_a).astype(np.int32)
        if step % 7:
            act[:, 0] = (step // 3) % 3; act[:, 1:] = (step // 3) % 2          # mostly a sane cyclic plan
        fp = rng.random((R, net.n_nodes, net.max_na), dtype=np.float32) if agent == "ma2c" else None
        fp_dev = None if fp is None else torch.from_numpy(fp).cuda()
        o1, r1, g1, d1 = gpu.step(torch.from_numpy(act).cuda(), fp_dev)
        o2, r2, g2, d2 = ref.step(act, fp)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(o1.cpu().numpy().view(np.uint32), o2.view(np.uint32))
        np.testing.assert_array_equal(r1.cpu().numpy().view(np.uint32), r2.view(np.uint32))
        np.testing.assert_array_equal(g1.cpu().numpy().view(np.uint32), g2.view(np.uint32))
        np.testing.assert_array_equal(d1.cpu().numpy(), d2)
        if step % 60 == 0 or step == 719:
            for r in (0, R - 1):
                c1, v1 = gpu.dump_state(r); c2, v2 = ref.dump_state(r)
                np.testing.assert_array_equal(c1, c2); np.testing.assert_array_equal(v1, v2)
    assert bool(d2.all()) and ref.misc(0)["arrived"] > 2000
    key = lambda t: t[np.lexsort(t.T[::-1])]
    for r in range(R):
        np.testing.assert_array_equal(key(gpu.trips(r)), key(ref.trips(r)))


def test_small_grid_env_reproduces_reference_trace():
    from deeprl_signal_control_b200.envs.small_grid_env import SmallGridController, SmallGridEnv
    z = np.load(os.path.join(GOLD, "small_greedy_test.npz"))
    meta = json.loads(str(z["meta"]))
    cp = configparser.ConfigParser()
    cp.read_string(SMALL_INI % "greedy")
    env = SmallGridEnv(cp["ENV_CONFIG"])
    assert env.n_s_ls == meta["n_s_ls"] and env.n_a_ls == meta["n_a_ls"] and env.node_names == meta["node_names"]
    env.train_mode = False
    ctrl = SmallGridController(env.node_names)
    ob = env.reset(test_ind=0)
    np.testing.assert_allclose(np.concatenate(ob), z["obs"][0], rtol=2e-6, atol=1e-6)
    for t in range(len(z["actions"])):
        assert [int(a) for a in ctrl.forward(ob)] == list(z["greedy"][t])
        ob, reward, done, greward = env.step(list(z["actions"][t]))
        np.testing.assert_allclose(np.concatenate(ob), z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(reward, z["reward"][t], rtol=3e-6, atol=1e-5)
        assert abs(greward - z["greward"][t]) < 1e-4

"""GPU: a scenario ingested from SUMO files (net/sumo_ingest.py) runs through the C ABI bit-exact against the
oracle, and the file-driven environment class (`envs/sumo_env.py:SumoNetEnv`) follows the reference's step protocol."""
import configparser
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "fixtures"))


def _mini(tmp_path):
    import make_mini_sumo
    return make_mini_sumo.write(str(tmp_path))


@pytest.mark.parametrize("agent", ["ma2c", "ia2c", "greedy"])
def test_ingested_scenario_bit_exact_vs_oracle(tmp_path, agent):
    from deeprl_signal_control_b200.net import sumo_ingest as ing
    from deeprl_signal_control_b200.net.tables import EnvParams
    from deeprl_signal_control_b200.sim import BatchedSim
    from oracle.sim_ref import RefSim
    net = ing.load_sumo_scenario(*_mini(tmp_path), agent=agent, use_wait=True)
    par = EnvParams(agent=agent, episode_length_sec=900)
    R = 7
    gpu, ref = BatchedSim(net, par, R), RefSim(net, par, R)
    seeds = np.arange(R, dtype=np.uint64) * np.uint64(7) + np.uint64(3)
    gpu.reset(seeds); ref.reset(seeds)
    rng = np.random.default_rng(11)
    na = np.array(net.n_a_ls)
    use_fp = agent == "ma2c"
    for step in range(180):
        act = (rng.integers(0, 1 << 20, (R, net.n_nodes)) % na).astype(np.int32)
        fp = rng.random((R, net.n_nodes, net.max_na), dtype=np.float32) if use_fp else None
        obs, rew, grew, done = gpu.step(torch.from_numpy(act).cuda(), torch.from_numpy(fp).cuda() if use_fp else None)
        o2, r2, g2, d2 = ref.step(act, fp)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(obs.cpu().numpy().view(np.uint32), o2.view(np.uint32))
        np.testing.assert_array_equal(rew.cpu().numpy().view(np.uint32), r2.view(np.uint32))
        np.testing.assert_array_equal(done.cpu().numpy(), d2)
        for a, b in zip(gpu.counts(), ref.counts()):
            np.testing.assert_array_equal(a.cpu().numpy(), b)
        if step % 45 == 0 or step == 179:
            c1, v1 = gpu.dump_state(R - 1); c2, v2 = ref.dump_state(R - 1)
            np.testing.assert_array_equal(c1, c2); np.testing.assert_array_equal(v1, v2)
    assert bool(done.cpu().numpy().all())


def test_file_driven_env_follows_the_reference_protocol(tmp_path):
    from deeprl_signal_control_b200.envs.sumo_env import SumoNetController, SumoNetEnv
    net_file, rou_file = _mini(tmp_path)
    cp = configparser.ConfigParser()
    cp.read_string("""
[ENV_CONFIG]
clip_wave = 2.0
clip_wait = 2.0
control_interval_sec = 5
agent = greedy
coop_gamma = 0.9
episode_length_sec = 900
norm_wave = 5.0
norm_wait = 100.0
coef_wait = 0.2
objective = hybrid
scenario = mini
seed = 5
test_seeds = 100,200
yellow_interval_sec = 2
net_file = %s
route_file = %s
""" % (net_file, rou_file))
    env = SumoNetEnv(cp["ENV_CONFIG"])
    env.train_mode = False
    env.init_test_seeds([100, 200])
    assert env.node_names == ["A", "B"] and env.n_a_ls == [2, 2]
    ctl = SumoNetController(env.node_names, env.nodes, {n: env.phase_map.phases[n].phases for n in env.node_names})
    ob = env.reset(test_ind=0)
    total, steps, done = 0.0, 0, False
    while not done:
        ob, reward, done, global_reward = env.step(ctl.forward(ob))
        assert len(ob) == 2 and all(np.isfinite(o).all() for o in ob)
        total += global_reward; steps += 1
    assert steps == 180 and total < 0
    env.terminate()

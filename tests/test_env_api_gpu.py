"""GPU: the reference-facing Python env (LargeGridEnv over libtsc) against the golden vectors
recorded from the reference's own envs/env.py (fake TraCI backed by the oracle)."""
import configparser
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

ENV_INI = """
[ENV_CONFIG]
clip_wave = 2.0
clip_wait = 2.0
control_interval_sec = 5
agent = %s
coop_gamma = 0.9
data_path = ./large_grid/data/
episode_length_sec = 3600
norm_wave = 5.0
norm_wait = 100.0
coef_wait = 0.2
peak_flow1 = 1100
peak_flow2 = 925
init_density = 0
objective = hybrid
scenario = large_grid
seed = 12
test_seeds = 10000,20000
yellow_interval_sec = 2
"""


@pytest.mark.parametrize("tag,agent,train", [("ma2c_train", "ma2c", True), ("ia2c_train", "ia2c", True),
                                             ("greedy_test", "greedy", False)])
def test_large_grid_env_reproduces_reference_trace(tag, agent, train):
    from deeprl_signal_control_b200.envs.large_grid_env import LargeGridEnv
    z = np.load(os.path.join(GOLD, "env_%s.npz" % tag))
    meta = json.loads(str(z["meta"]))
    cp = configparser.ConfigParser()
    cp.read_string(ENV_INI % agent)
    env = LargeGridEnv(cp["ENV_CONFIG"])
    env.train_mode = train
    assert env.n_s_ls == meta["n_s_ls"] and env.n_a_ls == meta["n_a_ls"]
    assert env.n_w_ls == meta["n_w_ls"] and env.n_f_ls == meta["n_f_ls"]
    assert env.node_names == meta["node_names"] and float(env.T) == meta["T"]
    ob = env.reset()
    assert isinstance(ob, list) and len(ob) == 25
    np.testing.assert_allclose(np.concatenate(ob), z["obs"][0], rtol=2e-6, atol=1e-6)
    for t in range(len(z["actions"])):
        if agent == "ma2c":
            env.update_fingerprint(list(z["fps"][t]))
        ob, reward, done, greward = env.step(list(z["actions"][t]))
        np.testing.assert_allclose(np.concatenate(ob), z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(np.asarray(reward) * np.ones(25), z["reward"][t], rtol=2e-6, atol=1e-5)
        assert abs(greward - z["greward"][t]) <= 1e-5 + 2e-6 * abs(z["greward"][t])
        assert done == bool(z["done"][t])
    env.terminate()

"""CPU: Monaco `real_net` ingest + oracle vs golden vectors from the reference's own RealNetEnv
(tests/golden/gen_real_net_golden.py)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def real_params(agent):
    from deeprl_signal_control_b200.net.tables import EnvParams
    return EnvParams(agent=agent, objective="queue", norm_wave=5.0, norm_wait=100.0, clip_wave=2.0, clip_wait=2.0,
                     coef_wait=0.0, coop_gamma=0.9, teleport_sec=300, real_net_norm=True, use_wait=False,
                     det_len=-1.0, halt_speed=0.1, queue_cap=10)


def test_structure_matches_survey_constants():
    """dims computed from most.net.xml + envs/real_net_env.py."""
    from deeprl_signal_control_b200.net.real_net import NODES, PHASES, real_net_tables
    net = real_net_tables("ma2c")
    assert net.n_s_ls == [32, 32, 17, 14, 14, 34, 6, 25, 19, 27, 37, 9, 37, 11, 21, 31, 6, 20, 21, 21, 27, 11, 23,
                          48, 24, 50, 34, 19]
    assert net.n_a_ls == [6, 4, 2, 2, 2, 4, 2, 4, 2, 5, 2, 2, 4, 2, 2, 4, 2, 3, 6, 3, 2, 4, 4, 4, 4, 4, 6, 3]
    assert net.n_det == 116 and net.n_nodes == 28 and net.n_routes == 16
    for name in net.node_names:        # phase string length == number of controlled links
        assert len(net.lanes_in[name]) == len(PHASES[NODES[name][0]][0])
    # asymmetric neighbour lists survive (envs/real_net_env.py:29,43)
    assert "cluster_9043_9052" in net.neighbor_map["9429"] and "9429" not in net.neighbor_map["cluster_9043_9052"]
    # every route is connected and ends with an arrival marker
    for r in range(net.n_routes):
        n = int(net.route_len[r])
        for h in range(n - 1):
            assert net.link_from[net.route_link[r, h]] == net.route_lane[r, h]
        assert net.route_link[r, n - 1] == -1


@pytest.mark.parametrize("tag,agent,train", [("ma2c_train", "ma2c", True), ("ia2c_train", "ia2c", True),
                                             ("greedy_test", "greedy", False)])
def test_real_net_step_matches_reference_python(tag, agent, train):
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    from oracle.sim_ref import RefSim
    z = np.load(os.path.join(GOLD, "real_%s.npz" % tag))
    meta = json.loads(str(z["meta"]))
    net, par = real_net_tables(agent), real_params(agent)
    assert net.node_names == meta["node_names"]
    assert net.n_s_ls == meta["n_s_ls"] and net.n_a_ls == meta["n_a_ls"]
    assert net.n_w_ls == meta["n_w_ls"] and net.n_f_ls == meta["n_f_ls"]
    for name in net.node_names:
        assert net.ilds_in[name] == meta["ilds_in"][name] and net.neighbor_map[name] == meta["neighbor"][name]
    sim = RefSim(net, par, 1)
    sim.reset([meta["seed0"]])
    sim.set_train_mode(train)
    fp0 = None
    if agent == "ma2c":
        fp0 = np.zeros((1, net.n_nodes, net.max_na), np.float32)
        for i, na in enumerate(net.n_a_ls):
            fp0[0, i, :na] = 1.0 / na
    np.testing.assert_allclose(sim.observe(fp0)[0], z["obs"][0], rtol=2e-6, atol=1e-6)
    for t in range(len(z["actions"])):
        fp = z["fps"][t][None] if agent == "ma2c" else None
        obs, rew, grew, done = sim.step(z["actions"][t][None], fp)
        np.testing.assert_allclose(obs[0], z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(rew[0], z["reward"][t], rtol=3e-6, atol=1e-5)
        np.testing.assert_allclose(grew[0], z["greward"][t], rtol=2e-6, atol=1e-5)
        assert bool(done[0]) == bool(z["done"][t])
    assert np.abs(z["greward"]).max() >= 10


def test_real_net_greedy_controller_matches_reference():
    from deeprl_signal_control_b200.envs.env import Node
    from deeprl_signal_control_b200.envs.real_net_env import RealNetController
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    z = np.load(os.path.join(GOLD, "real_greedy_test.npz"))
    net = real_net_tables("greedy")
    nodes = {}
    for name in net.node_names:
        nd = Node(name)
        nd.lanes_in, nd.ilds_in = net.lanes_in[name], net.ilds_in[name]
        nodes[name] = nd
    ctrl = RealNetController(net.node_names, nodes)
    off = net.node_obs_off
    for t in range(len(z["greedy"])):
        ob = [z["obs"][t][off[i]:off[i + 1]] for i in range(net.n_nodes)]
        assert list(ctrl.forward(ob)) == list(z["greedy"][t])


def test_recorded_ma2c_plans_replay_quantifies_the_gap_to_sumo():
    """Distribution-level comparison of the restated dynamics with SUMO, at the vType that SHIPS (tau = 1.0, nothing
    overridden): replay the signal plans the reference RECORDED for its trained MA2C agent on Monaco
    (real_net_experimental_data/eva_data, 10 evaluation episodes; fixture cut by tests/golden/replay_monaco_eval_traces.py)
    through the oracle, open loop, and compare with what SUMO produced under the same plans.

    This is a CHARACTERISATION of a known gap, not a parity claim: the recordings were made with the
    paper-era vType (tau = 0.5 on SUMO 0.32, reference README.md:63); with a 1-s Euler step the Krauss stop speed
    overshoots for tau < 1 s (SUMO then makes "emergency stops" at the lane end, and the reference removed tau = 0.5
    because of collisions), so that regime cannot be restated faithfully, and at tau = 1.0 the recorded plans meet 30-40 %
    less junction capacity than they were made for: the network spills back.  The bounds below pin the size of that gap
    (and the demand side, which must agree): a change of the model that moves them is a change of fidelity."""
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    from oracle.sim_ref import RefSim
    z = np.load(os.path.join(GOLD, "monaco_ma2c_recorded_traces.npz"))
    want = json.loads(str(z["recorded"]))
    acts = z["actions"].astype(np.int32)                         # [episodes, 720, 28]
    net, par = real_net_tables("greedy"), real_params("greedy")
    assert par.tau == 1.0
    R = acts.shape[0]
    sim = RefSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64) + np.uint64(10000))
    sim.set_train_mode(False)
    sim.set_record(True)
    peak, speed = np.zeros(R), []
    for t in range(acts.shape[1]):
        st = sim.step_record(acts[:, t])[4]
        peak = np.maximum(peak, st[..., 0].max(1))
        speed.append(st[..., 4].mean())
    trips = np.concatenate([sim.trips(r) for r in range(R)])
    departed = np.mean([sim.misc(r)["departed"] for r in range(R)])
    backlog = np.mean([sim.misc(r)["backlog"] for r in range(R)])
    ratio = lambda got, key: got / want[key]
    # demand side: everything the flows generate is either in the network, arrived, or waiting to be inserted
    assert abs((departed + backlog) - 2464.0) < 1.0 and want["departed_per_episode"] <= 2464.0
    # supply side: the documented deficit (SUMO completed 2397 trips per episode under these plans)
    assert 0.35 < ratio(len(trips) / R, "trips_per_episode") < 0.60
    assert 0.30 < ratio(float(np.mean(speed)), "avg_speed_mps") < 0.60
    assert 1.8 < ratio(float(peak.mean()), "peak_cars") < 3.2

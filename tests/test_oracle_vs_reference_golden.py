"""CPU tests: the oracle's control protocol / observation / reward / shaping and the host-side
scenario tables, checked against golden vectors produced by the REFERENCE's own Python
(tests/golden/gen_env_golden.py: envs/env.py + envs/large_grid_env.py + build_file.py executed
over a fake TraCI connection).  Tolerances: integer / string outputs exact; float outputs are
f32 here vs f64 in the reference -> rtol 2e-6, atol 1e-6 (stated per assert).
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(tag):
    z = np.load(os.path.join(GOLD, "env_%s.npz" % tag), allow_pickle=False)
    return z, json.loads(str(z["meta"]))


def _mask_to_str(open_m, major_m, yel_m, n=12):
    return "".join("y" if (yel_m >> i) & 1 else ("G" if (major_m >> i) & 1 else ("g" if (open_m >> i) & 1 else "r"))
                   for i in range(n))


def _phase_strings(net, prev, act, yellow):
    """Python mirror of node_signal() (oracle/tsc_sim_ref.c, csrc/tsc_sim.cu)."""
    out = []
    for i in range(net.n_nodes):
        g1, m1 = int(net.node_green[i, act[i]]), int(net.node_major[i, act[i]])
        o, m, y = g1, m1, 0
        if yellow and prev[i] != act[i]:
            g0 = int(net.node_green[i, prev[i]])
            sw_red, sw_green = g0 & ~g1, ~g0 & g1
            if sw_red:
                y, o, m = sw_red, g1 & ~sw_green, m1 & ~sw_green
        out.append(_mask_to_str(o, m, y))
    return out


@pytest.mark.parametrize("tag,agent,train", [("ma2c_train", "ma2c", True), ("ia2c_train", "ia2c", True),
                                             ("greedy_test", "greedy", False), ("ma2c_test", "ma2c", False)])
def test_env_step_matches_reference_python(tag, agent, train):
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    z, meta = _load(tag)
    net, par = build_large_grid(agent=agent), EnvParams(agent=agent)
    # static structure (envs/env.py:207-254,303-323)
    assert net.node_names == meta["node_names"]
    assert net.n_s_ls == meta["n_s_ls"] and net.n_a_ls == meta["n_a_ls"]
    assert net.n_w_ls == meta["n_w_ls"] and net.n_f_ls == meta["n_f_ls"]
    for name in net.node_names:
        assert net.ilds_in[name] == meta["ilds_in"][name]
        assert net.neighbor_map[name] == meta["neighbor"][name]
    sim = RefSim(net, par, 1)
    sim.reset([meta["seed0"]])
    sim.set_train_mode(train)
    # reset() installs uniform fingerprints first (envs/env.py:556-557,263-269)
    fp0 = np.full((1, net.n_nodes, net.max_na), 1.0 / 5, np.float32) if agent == "ma2c" else None
    np.testing.assert_allclose(sim.observe(fp0)[0], z["obs"][0], rtol=2e-6, atol=1e-6)
    prev = np.zeros(net.n_nodes, np.int64)
    for t in range(len(z["actions"])):
        act = z["actions"][t]
        fp = z["fps"][t][None] if agent == "ma2c" else None
        # phase strings the reference sent to SUMO (envs/env.py:128-152,455-459): exact
        assert _phase_strings(net, prev, act, True) == list(z["yellow"][t])
        assert _phase_strings(net, prev, act, False) == list(z["green"][t])
        prev = act.astype(np.int64)
        obs, rew, grew, done = sim.step(act[None], fp)
        np.testing.assert_allclose(obs[0], z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(rew[0], z["reward"][t], rtol=2e-6, atol=1e-5)
        np.testing.assert_allclose(grew[0], z["greward"][t], rtol=2e-6, atol=1e-5)
        assert bool(done[0]) == bool(z["done"][t])
        assert np.array_equal(sim.counts()[3][0], act)          # phase index applied: exact
    assert np.abs(z["greward"]).max() > 20                      # the trace carried real traffic


def test_greedy_controller_matches_reference():
    """LargeGridController.greedy (envs/large_grid_env.py:56-60) on the recorded observations."""
    from deeprl_signal_control_b200.envs.large_grid_env import LargeGridController
    z, meta = _load("greedy_test")
    ctrl = LargeGridController(meta["node_names"])
    for t in range(len(z["greedy"])):
        ob = z["obs"][t].reshape(25, 6)
        assert list(ctrl.forward(list(ob))) == list(z["greedy"][t])


def test_grid_tables_match_reference_generator():
    """Edges / connections / detectors / flows vs the XML strings the reference emits
    (large_grid/data/build_file.py:66-124, 268-326, 360-391)."""
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    spec = json.load(open(os.path.join(GOLD, "grid_spec.json")))
    net = build_large_grid(agent="ma2c")
    # edges: same ids, same order, lanes per type (a = 2 lanes, b = 1)
    lanes = []
    for eid, a, b, typ in spec["edges"]:
        assert eid == "%s_%s" % (a, b)
        lanes += ["%s_%d" % (eid, k) for k in range(2 if typ == "a" else 1)]
    assert lanes == net.lane_names
    # lane lengths = Euclidean node distance from the reference's node coordinates
    xy = {n: (float(x), float(y)) for n, x, y, _ in spec["nodes"]}
    for k, name in enumerate(net.lane_names):
        a, b, _ = name.split("_")
        d = np.hypot(xy[a][0] - xy[b][0], xy[a][1] - xy[b][1])
        assert abs(d - net.lane_len[k]) < 1e-4
    # connections: the reference's 300 (fromLane -> toLane) pairs == our link table
    ref = sorted((f + "_" + fl, t + "_" + tl) for f, t, fl, tl in spec["connections"])
    ours = sorted((net.lane_names[net.link_from[k]], net.lane_names[net.link_to[k]]) for k in range(net.n_links))
    assert ref == ours
    # detectors: one per incoming lane of a signalised node == union of our ilds_in
    assert sorted(spec["ilds"]) == sorted(net.lane_names[l] for l in net.det_lane)
    # flows: (from edge, to edge, begin, end, vehsPerHour) multiset
    ref_f = sorted((fr, to, int(b), int(e), int(q)) for _, fr, to, b, e, q in spec["flows"])
    ours_f = []
    for s, b, e, q in net.flow_list:
        r = int(net.src_route[s])
        first = net.lane_names[net.route_lane[r, 0]].rsplit("_", 1)[0]
        last = net.lane_names[net.route_lane[r, net.route_len[r] - 1]].rsplit("_", 1)[0]
        ours_f.append((first, last, b, e, q))
    assert ref_f == sorted(ours_f)
    # every vehicle of every flow becomes due exactly once
    assert int(net.src_due.sum()) == sum(-(-(e - b) * q // 3600) for _, _, b, e, q in ref_f)


def test_routes_are_connected_and_use_turn_lanes():
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    net = build_large_grid()
    for r in range(net.n_routes):
        n = int(net.route_len[r])
        for h in range(n - 1):
            k = int(net.route_link[r, h])
            assert net.link_from[k] == net.route_lane[r, h]
            # destination edge of the link contains the next hop's lane
            a = net.lane_names[net.link_to[k]].rsplit("_", 1)[0]
            b = net.lane_names[net.route_lane[r, h + 1]].rsplit("_", 1)[0]
            assert a == b
        assert net.route_link[r, n - 1] == -1


def test_initial_fleet_matches_reference_generator():
    """init_density > 0: the 120 (edge, departLane, sink, number) groups of `init_routes`
    (large_grid/data/build_file.py:223-266), including the sink edges the reference draws from numpy's global generator
    after np.random.seed(seed), and their effect on the tables (all due in second 0, routes end on the drawn sink)."""
    from deeprl_signal_control_b200.net.large_grid import build_large_grid, init_fleet_specs
    spec = json.load(open(os.path.join(GOLD, "grid_spec.json")))
    for seed in (12, 31):
        ref = [(fr, to, int(lane), int(num)) for _, fr, to, lane, num in spec["init_fleet"][str(seed)]]
        ours = [("%s_%s" % (a, b), "%s_%s" % sink, lane, n) for a, b, lane, sink, n in init_fleet_specs(0.2, seed)]
        assert ours == ref and len(ref) == 120
    net = build_large_grid(agent="ma2c", init_density=0.2, seed=12)
    base = build_large_grid(agent="ma2c")
    assert net.n_routes == base.n_routes + 120 and net.n_src == base.n_src + 120
    assert int(net.src_due[0].sum()) - int(base.src_due[0].sum()) == 120 * 6        # int(30 * 0.2) per group
    ref = spec["init_fleet"]["12"]
    for k in range(120):
        r = base.n_routes + k
        first = net.lane_names[net.route_lane[r, 0]].rsplit("_", 1)[0]
        last = net.lane_names[net.route_lane[r, net.route_len[r] - 1]].rsplit("_", 1)[0]
        assert (first, last) == (ref[k][1], ref[k][2])
        for h in range(int(net.route_len[r]) - 1):                                   # connected
            assert net.link_from[net.route_link[r, h]] == net.route_lane[r, h]

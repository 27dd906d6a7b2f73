"""CPU: 6-intersection `small_grid` scenario — tables vs the reference generator's constants, and the oracle vs
golden vectors from the reference's own SmallGridEnv (tests/golden/gen_small_grid_golden.py: envs/env.py +
envs/small_grid_env.py over a fake TraCI connection).  Float tolerances as in test_oracle_vs_reference_golden.py."""
import json
import os

import numpy as np

from tests.test_oracle_vs_reference_golden import _mask_to_str

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _params(meta, agent="greedy"):
    from deeprl_signal_control_b200.net.tables import EnvParams
    c = meta["cfg"]
    return EnvParams(agent=agent, objective=c["objective"], norm_wave=float(c["norm_wave"]),
                     norm_wait=float(c["norm_wait"]), clip_wave=float(c["clip_wave"]), clip_wait=float(c["clip_wait"]),
                     coef_wait=float(c["coef_wait"]), coop_gamma=float(c["coop_gamma"]))


def _phase_strings(net, prev, act, yellow):
    out = []
    for i in range(net.n_nodes):
        n = len(net.lanes_in[net.node_names[i]])
        g1, m1 = int(net.node_green[i, act[i]]), int(net.node_major[i, act[i]])
        o, m, y = g1, m1, 0
        if yellow and prev[i] != act[i]:
            g0 = int(net.node_green[i, prev[i]])
            sw_red, sw_green = g0 & ~g1, ~g0 & g1
            if sw_red:
                y, o, m = sw_red, g1 & ~sw_green, m1 & ~sw_green
        out.append(_mask_to_str(o, m, y, n))
    return out


def test_tables_match_generator_constants():
    """small_grid/data/build_file.py: 20 one-lane edges, connection list, detectors on the 14 non-sink lanes,
    turn ratios; envs/small_grid_env.py: 3 phases at nt1, 2 elsewhere."""
    from deeprl_signal_control_b200.net.small_grid import build_small_grid, small_connections
    net = build_small_grid()
    assert net.n_lanes == 20 and net.n_links == len(small_connections()) == 31 and net.n_nodes == 6
    assert net.n_a_ls == [3, 2, 2, 2, 2, 2] and net.n_det == 13
    assert (net.link_node >= 0).sum() == 29                      # the two links of the priority junction npc are free
    assert abs(float(net.lane_len[net.lane_names.index("nt1_npc_0")]) - 200.0) < 1e-3
    for name in net.node_names:
        assert len(net.phases[name][0]) == len(net.lanes_in[name])
    # route choice of every origin is a partition of [0, 1) in every 10-minute interval; the Bernoulli flows keep p
    for g in range(5):
        q = np.where(net.src_group == g)[0]
        np.testing.assert_allclose(net.src_plo[:, q[0]], 0.0)
        np.testing.assert_allclose(net.src_phi[:, q[-1]], 1.0)
        np.testing.assert_allclose(net.src_phi[:, q[:-1]], net.src_plo[:, q[1:]])
        assert (np.diff(net.src_due[:, q], axis=1) == 0).all()   # siblings share the origin's due times
    mf = np.where(net.src_group >= 5)[0]
    assert len(mf) == 10 and np.allclose(net.src_phi[:, mf], 0.28) and np.allclose(net.src_plo[:, mf], 0.0)
    # np1 -> {nt2 0.2, nt6 0.5, npc 0.3}, np9 -> {nt3 0.6, nt5 0.4}  (build_file.py:223-243)
    q1 = np.where(net.src_group == 0)[0]
    names = [net.route_names[net.src_route[q]] for q in q1]
    p = net.src_phi[0, q1] - net.src_plo[0, q1]
    assert abs(sum(x for x, n in zip(p, names) if "nt1_nt2" in n) - 0.2) < 1e-6
    assert abs(sum(x for x, n in zip(p, names) if "nt1_nt6" in n) - 0.5) < 1e-6
    assert abs(sum(x for x, n in zip(p, names) if "nt1_npc" in n) - 0.3) < 1e-6
    for r in range(net.n_routes):                                 # connected, end with an arrival marker
        n = int(net.route_len[r])
        for h in range(n - 1):
            lk = net.route_link[r, h]
            assert net.link_from[lk] == net.route_lane[r, h] and net.link_to[lk] == net.route_lane[r, h + 1]
        assert net.route_link[r, n - 1] == -1


def test_small_grid_step_matches_reference_python():
    from deeprl_signal_control_b200.envs.small_grid_env import SmallGridController
    from deeprl_signal_control_b200.net.small_grid import build_small_grid
    from oracle.sim_ref import RefSim
    z = np.load(os.path.join(GOLD, "small_greedy_test.npz"))
    meta = json.loads(str(z["meta"]))
    net = build_small_grid(int(meta["cfg"]["num_extra_car_per_hour"]), agent="greedy",
                           coop_gamma=float(meta["cfg"]["coop_gamma"]))
    par = _params(meta)
    assert net.node_names == meta["node_names"]
    assert net.n_s_ls == meta["n_s_ls"] and net.n_a_ls == meta["n_a_ls"] and net.n_w_ls == meta["n_w_ls"]
    for name in net.node_names:
        assert net.ilds_in[name] == meta["ilds_in"][name]
    sim = RefSim(net, par, 1)
    sim.reset([meta["seed0"]])
    sim.set_train_mode(False)
    np.testing.assert_allclose(sim.observe()[0], z["obs"][0], rtol=2e-6, atol=1e-6)
    ctrl = SmallGridController(net.node_names)
    off = net.node_obs_off
    prev = np.zeros(net.n_nodes, np.int64)
    for t in range(len(z["actions"])):
        ob = [z["obs"][t][off[i]:off[i + 1]] for i in range(net.n_nodes)]
        assert [int(a) for a in ctrl.forward(ob)] == list(z["greedy"][t])     # envs/small_grid_env.py:52-57
        act = z["actions"][t]
        assert _phase_strings(net, prev, act, True) == list(meta["yellow"][t])
        assert _phase_strings(net, prev, act, False) == list(meta["green"][t])
        prev = act.astype(np.int64)
        obs, rew, grew, done = sim.step(act[None])
        np.testing.assert_allclose(obs[0], z["obs"][t + 1], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(rew[0], z["reward"][t], rtol=3e-6, atol=1e-5)
        np.testing.assert_allclose(grew[0], z["greward"][t], rtol=2e-6, atol=1e-5)
        assert bool(done[0]) == bool(z["done"][t])
    assert np.abs(z["greward"]).max() > 50


def _mix32(h):
    h &= 0xffffffff
    h ^= h >> 16; h = (h * 0x7feb352d) & 0xffffffff
    h ^= h >> 15; h = (h * 0x846ca68b) & 0xffffffff
    h ^= h >> 16
    return h


def _rng_u32(s0, s1, a, b, c):
    """Python mirror of rng_draw(rng_key(s0, s1, a), b, c) (oracle/tsc_sim_ref.c, csrc/tsc_sim.cu): two rounds for the
    key of (replica seed, second), one round per draw."""
    key = _mix32(_mix32(s0 ^ ((a * 0x9E3779B1) & 0xffffffff)) ^ s1)
    return _mix32(key ^ ((b * 0x85EBCA77 + c * 0xC2B2AE3D) & 0xffffffff))


def test_stochastic_demand_accounting_is_exact():
    """Every due vehicle is either inserted or still pending: per source, (departed by route) + (backlog) equals the
    number of draws that fell into the source's probability interval, recomputed here from the counter RNG
    (include/tsc.h: u per (replica seed, second, group); keep iff plo <= u < phi of the 10-minute interval)."""
    from deeprl_signal_control_b200.net.small_grid import build_small_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    net = build_small_grid()
    par = EnvParams(agent="greedy", norm_wave=1.0, norm_wait=1.0, clip_wave=1000.0, clip_wait=1000.0)
    seeds = np.array([7, (5 << 32) + 9], np.uint64)
    sim = RefSim(net, par, 2)
    sim.reset(seeds)
    sim.set_record(True)
    act = np.zeros((2, net.n_nodes), np.int32)
    T = 360
    for t in range(T):
        act[:, 0] = (t // 3) % 3; act[:, 1:] = (t // 3) % 2
        sim.step(act)
    for r in range(2):
        routes = np.concatenate([sim.trips(r)[:, 2], (sim.dump_state(r)[1][:, 2] >> 16) & 255])
        dep_by_route = np.bincount(routes, minlength=net.n_routes)
        got = dep_by_route[net.src_route] + sim.backlog(r)
        s_lo, s_hi = int(seeds[r]) & 0xffffffff, int(seeds[r]) >> 32
        exp = np.zeros(net.n_src, np.int64)
        for t in range(T * par.control_interval_sec):
            iv = min(t // net.pint_sec, net.n_pint - 1)
            for g in np.unique(net.src_group):
                u = np.float32((_rng_u32(s_lo, s_hi, t, 0x20000 + int(g), 7) >> 8) * (1.0 / 16777216.0))
                for q in np.where(net.src_group == g)[0]:
                    if net.src_due[t, q] and net.src_plo[iv, q] <= u < net.src_phi[iv, q]:
                        exp[q] += int(net.src_due[t, q])
        np.testing.assert_array_equal(got, exp)
        assert exp.sum() > 1500


def test_stochastic_demand_statistics():
    """Route shares and Bernoulli rates realised by the counter-RNG draws match the tables (law of large numbers over
    replicas).  Shares are measured on the departed vehicles of a lightly loaded network (no extra flows): origin np1
    sends 20 / 50 / 30 % towards nt2 / nt6 / npc.  The `probability=0.28` flows
    are measured on departures per second in the first minutes, before their entry lanes saturate."""
    from deeprl_signal_control_b200.net.small_grid import build_small_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    par = EnvParams(agent="greedy", norm_wave=1.0, norm_wait=1.0, clip_wave=1000.0, clip_wait=1000.0)
    R = 16
    net = build_small_grid(num_car_hourly=0)
    sim = RefSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64) + np.uint64(100))
    sim.set_record(True)
    act = np.zeros((R, net.n_nodes), np.int32)
    for t in range(720):
        act[:, 0] = (t // 3) % 3; act[:, 1:] = (t // 3) % 2
        sim.step(act)
    # routes of every departed vehicle = arrived trips + vehicles still in the network (route id in the state word)
    routes = np.concatenate([sim.trips(r)[:, 2] for r in range(R)] +
                            [(sim.dump_state(r)[1][:, 2] >> 16) & 255 for r in range(R)])
    assert len(routes) == sum(sim.misc(r)["departed"] for r in range(R))
    names = np.array(net.route_names)[routes]
    from_np1 = np.array([n.startswith("np1_nt1") for n in names])
    share = lambda key: np.mean([key in n for n in names[from_np1]])
    assert from_np1.sum() > 3000
    # (the vehicles still pending at the end are spread evenly over the six sibling routes, not by their ratios)
    assert abs(share("nt1_nt2") - 0.2) < 0.03 and abs(share("nt1_nt6") - 0.5) < 0.06 and abs(share("nt1_npc") - 0.3) < 0.07
    # extra flows 3, 4, 5 run during the first 20 minutes (build_file.py:184,200-207), 0.28 vehicles per second each
    net = build_small_grid(num_car_hourly=1000)
    sim = RefSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64) + np.uint64(300))
    dep0 = np.array([sim.misc(r)["departed"] for r in range(R)])
    for t in range(24):                                   # 120 s
        act[:, 0] = t % 3; act[:, 1:] = t % 2
        sim.step(act)
    dep = np.array([sim.misc(r)["departed"] + sim.misc(r)["backlog"] for r in range(R)]) - dep0
    jtr = sum(net.src_due[:120, np.where(net.src_group == g)[0][0]].sum() for g in range(5))     # deterministic part
    assert abs((dep.mean() - jtr) / (3 * 120.0) - 0.28) < 0.03

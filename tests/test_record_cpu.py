"""CPU: the evaluation / recording path of the oracle (per-second traffic statistics of
`_measure_traffic_step`, envs/env.py:409-437, and the tripinfo rows the reference reads from SUMO's
--tripinfo-output, envs/env.py:498-515)."""
import numpy as np
import pytest


@pytest.mark.parametrize("scenario", ["grid", "monaco"])
def test_step_record_equals_step_and_stats_are_consistent(scenario):
    from oracle.sim_ref import RefSim
    if scenario == "grid":
        from deeprl_signal_control_b200.net.large_grid import build_large_grid
        from deeprl_signal_control_b200.net.tables import EnvParams
        net, par = build_large_grid(agent="ia2c"), EnvParams(agent="ia2c")
    else:
        from deeprl_signal_control_b200.net.real_net import real_net_tables
        from tests.test_real_net_cpu import real_params
        net, par = real_net_tables("ia2c"), real_params("ia2c")
    R, n_steps = 2, 150
    a, b = RefSim(net, par, R), RefSim(net, par, R)
    seeds = np.array([5, 9], np.uint64)
    a.reset(seeds); b.reset(seeds)
    b.set_record(True)
    rng = np.random.default_rng(0)
    n_a = np.asarray(net.n_a_ls)
    prev = np.zeros((R, 8), np.float32)
    for t in range(n_steps):
        act = (rng.integers(0, 1 << 30, size=(R, net.n_nodes)) % n_a).astype(np.int32)
        o1, r1, g1, d1 = a.step(act)
        o2, r2, g2, d2, st = b.step_record(act)
        np.testing.assert_array_equal(o1.view(np.uint32), o2.view(np.uint32))
        np.testing.assert_array_equal(r1.view(np.uint32), r2.view(np.uint32))
        np.testing.assert_array_equal(g1.view(np.uint32), g2.view(np.uint32))
        assert st.shape == (R, par.control_interval_sec, 8)
        # live = departed - arrived after every second; cumulative counters never decrease
        np.testing.assert_array_equal(st[..., 0], st[..., 1] - st[..., 2])
        assert (np.diff(np.concatenate([prev[:, None, 1:3], st[..., 1:3]], 1), axis=1) >= 0).all()
        assert (st[..., 3] >= 0).all() and (st[..., 4] >= 0).all() and (st[..., 6] >= 0).all()
        prev = st[:, -1]
    np.testing.assert_array_equal(b.traffic_stats(), prev)          # last second == statistics of the final state
    for r in range(R):
        c1, v1 = a.dump_state(r); c2, v2 = b.dump_state(r)
        np.testing.assert_array_equal(c1, c2); np.testing.assert_array_equal(v1, v2)
        trips = b.trips(r)
        m = b.misc(r)
        assert len(trips) == m["arrived"] > 0
        dep, arr, route, wsec, wcnt = trips.T
        assert (arr > dep).all() and (arr <= m["cur_sec"]).all()
        assert (wsec <= arr - dep).all() and (wcnt <= wsec).all() and ((wcnt > 0) == (wsec > 0)).all()
        assert len({(int(x), int(y)) for x, y in zip(route, dep)}) == len(trips)      # (route, depart) is a unique id
        assert (np.diff(arr) >= 0).all()                                              # arrival order


def test_record_mode_does_not_change_the_hot_state():
    """Record mode only adds the trip words: a recorded and an unrecorded replica stay identical."""
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    net, par = build_large_grid(agent="greedy"), EnvParams(agent="greedy")
    a, b = RefSim(net, par, 1), RefSim(net, par, 1)
    a.reset(np.array([3], np.uint64)); b.reset(np.array([3], np.uint64))
    b.set_record(True)
    act = np.zeros((1, net.n_nodes), np.int32)
    for t in range(120):
        act[:] = (t // 6) % 5
        o1 = a.step(act)[0]; o2 = b.step(act)[0]          # plain step() with record mode on
        np.testing.assert_array_equal(o1, o2)
    assert len(b.trips(0)) == b.misc(0)["arrived"]


def test_multi_step_run_equals_stepping():
    """ref_run_mt (one thread pool for n steps: the CPU-baseline driver of bench.py) == n calls of ref_step_mt."""
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    net, par = build_large_grid(agent="ma2c"), EnvParams(agent="ma2c")
    R = 6
    a, b = RefSim(net, par, R), RefSim(net, par, R)
    seeds = np.arange(R, dtype=np.uint64) + np.uint64(40)
    a.reset(seeds); b.reset(seeds)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, (4, R, net.n_nodes), dtype=np.int32)
    fp = rng.random((R, net.n_nodes, net.max_na), dtype=np.float32)
    for t in range(30):
        oa = a.step(acts[t % 4], fp, threads=3)
    ob = b.run(acts, 30, fp, threads=4)
    for x, y in zip(oa, ob):
        np.testing.assert_array_equal(x, y)
    for r in range(R):
        np.testing.assert_array_equal(a.dump_state(r)[1], b.dump_state(r)[1])

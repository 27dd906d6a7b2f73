"""CPU: size-independent invariants of the simulation model on the oracle (the CUDA kernel is bit-identical to it, so they
transfer): vehicle conservation, collision freedom inside a lane (FIFO order with at least the vehicle length between
fronts), speed and waiting-time bounds, ring capacities, route consistency, and linearity of the n-step returns."""
import numpy as np
import pytest


def _scenarios():
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    from deeprl_signal_control_b200.net.small_grid import build_small_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    from tests.test_real_net_cpu import real_params
    return {"grid": (build_large_grid(agent="ia2c"), EnvParams(agent="ia2c")),
            "monaco": (real_net_tables("ia2c"), real_params("ia2c")),
            "small": (build_small_grid(agent="greedy"),
                      EnvParams(agent="greedy", norm_wave=1.0, norm_wait=1.0, clip_wave=1000.0, clip_wait=1000.0))}


@pytest.mark.parametrize("name", ["grid", "monaco", "small"])
def test_state_invariants_under_random_control(name):
    from oracle.sim_ref import RefSim
    net, par = _scenarios()[name]
    R = 3
    sim = RefSim(net, par, R)
    sim.reset(np.array([21, 22, 23], np.uint64))
    rng = np.random.default_rng(6)
    n_a = np.asarray(net.n_a_ls)
    vmax_sf = 0.5 + 255.0 / 256.0                         # largest speed factor the 8-bit field can hold
    for step in range(360):
        if step % 120 < 80:
            act = (rng.integers(0, 1 << 30, size=(R, net.n_nodes)) % n_a).astype(np.int32)
        else:
            act = np.zeros((R, net.n_nodes), np.int32)       # hold one phase: queues build up
        sim.step(act)
        if step % 12:
            continue
        for r in range(R):
            cnt, veh = sim.dump_state(r)
            m = sim.misc(r)
            assert cnt.sum() == m["live"] == m["departed"] - m["arrived"]
            assert (cnt <= net.lane_cap).all() and (cnt >= 0).all()
            pos = veh[:, 0].copy().view(np.float32); spd = veh[:, 1].copy().view(np.float32)
            wait = veh[:, 2] & 1023; hop = (veh[:, 2] >> 10) & 63; route = (veh[:, 2] >> 16) & 255
            assert np.isfinite(pos).all() and np.isfinite(spd).all()
            assert (spd >= 0).all() and (wait <= min(1023, m["cur_sec"])).all()
            assert (route < net.n_routes).all() and (hop < net.route_len[route]).all()
            k = 0
            for l, c in enumerate(cnt):
                if c == 0:
                    continue
                p, v = pos[k:k + c], spd[k:k + c]
                # the vehicle is on the lane its route says, inside the lane, and never faster than the fastest lane
                # allows (it may exceed its CURRENT lane's limit for the second in which it entered from a faster one)
                assert (net.route_lane[route[k:k + c], hop[k:k + c]] == l).all()
                # (a vehicle may enter a lane shorter than one second of travel beyond its end — Monaco has 6 m lanes —
                #  and then crosses on in the next second: at most one lane per second)
                assert (p >= 0).all() and (p < net.lane_len[l] + net.lane_vmax.max() * vmax_sf).all()
                assert (v <= net.lane_vmax.max() * vmax_sf + 1e-3).all()
                # FIFO order, front vehicle first; fronts at least one vehicle length apart (no overlap)
                assert (np.diff(p) <= -(par.veh_len - 1e-3)).all(), (name, step, l, p)
                k += c
            # a vehicle that moved has no waiting time
            assert (wait[spd >= 0.1] == 0).all()


def test_nstep_returns_are_linear_and_match_closed_form():
    from oracle.learner_ref import nstep_returns
    rng = np.random.default_rng(0)
    T, gamma = 9, 0.9
    r1, r2 = rng.normal(size=(T, 4)), rng.normal(size=(T, 4))
    v = rng.normal(size=(T, 4))
    dones = [0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0]
    b1, b2 = rng.normal(size=4), rng.normal(size=4)
    R1, A1 = nstep_returns(list(r1), list(v), dones, b1, gamma)
    R2, _ = nstep_returns(list(r2), list(v), dones, b2, gamma)
    R12, A12 = nstep_returns(list(2 * r1 - r2), list(v), dones, 2 * b1 - b2, gamma)
    np.testing.assert_allclose(R12, 2 * R1 - R2, atol=1e-12)
    np.testing.assert_allclose(A12, R12 - v, atol=1e-12)
    # closed form of the last segment (after the last done): R_t = sum_k gamma^k r_{t+k} + gamma^(T-t) boot
    np.testing.assert_allclose(R1[8], r1[8] + gamma * b1, atol=1e-12)
    # a `done` cuts the bootstrap: R_7 = r_7 exactly
    np.testing.assert_allclose(R1[7], r1[7], atol=1e-12)


def test_queue_discharge_headway_at_a_signal(tmp_path):
    """Model-level check that needs no reference numbers: a standing queue released by a green signal discharges with
    the Krauss car-following headway tau + (length + minGap) / v, i.e. one vehicle per ~2 s (about 1800 veh/h per lane)
    for the reference's vType (accel 5, decel 10, sigma 0.5, tau 1, length 5, minGap 2.5) — the saturation flow SUMO users
    quote for tau = 1.  (The gap to the reference's recorded congestion, DESIGN.md 2.2, is therefore not a stop-line
    capacity error of a factor.)"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures"))
    import make_mini_sumo
    from deeprl_signal_control_b200.net import sumo_ingest as ing
    from deeprl_signal_control_b200.net.tables import EnvParams
    from oracle.sim_ref import RefSim
    old = make_mini_sumo.L_EDGE
    make_mini_sumo.L_EDGE = 900.0                      # room for a 120-vehicle queue
    try:
        net_file, rou_file = make_mini_sumo.write(str(tmp_path))
    finally:
        make_mini_sumo.L_EDGE = old
    with open(rou_file, "w") as f:
        f.write('<routes>\n <vType id="car" length="5" accel="5" decel="10"/>\n <flow id="f0" from="W_A" to="B_E" via="A_B" '
                'begin="0" end="3600" vehsPerHour="3600" type="car"/>\n</routes>\n')
    net = ing.load_sumo_scenario(net_file, rou_file, agent="greedy", use_wait=True)
    par = EnvParams(agent="greedy", episode_length_sec=3600, control_interval_sec=1, yellow_interval_sec=0)
    sim = RefSim(net, par, 1)
    sim.reset(np.array([3], np.uint64)); sim.set_train_mode(False)
    l_ab = [i for i, nm in enumerate(net.lane_names) if nm.startswith("A_B")][0]
    crossed = []
    for sec in range(700):
        sim.step(np.array([[0 if sec < 600 else 1, 1]], np.int32))       # west approach red for 600 s, then green
        if sec >= 600:
            cnt, _ = sim.dump_state(0)
            crossed.append(int(cnt[l_ab]))              # nobody leaves A_B (909 m) within these 100 s
    assert sim.misc(0)["live"] > 100                    # the queue was there
    n20, n90 = crossed[20], crossed[90]
    headway = 70.0 / (n90 - n20)
    assert 1.6 <= headway <= 2.4, headway               # ~1800 veh/h per lane

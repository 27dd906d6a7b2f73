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

"""CPU: properties of the restated Krauss car-following helpers (oracle/tsc_sim_ref.c; the CUDA kernel carries the same
IEEE-binary32 operation sequence).  SUMO itself is absent, so these pin what the formulas must guarantee instead of
golden numbers: a vehicle that takes `stop_speed(gap)` now and then brakes with `b` per second (Euler, dt = 1 s, reaction
time tau) never travels farther than `gap`; speeds are monotone in the gap; following at `follow_speed` keeps the
vehicle behind a leader that brakes as hard as it can; `free_speed` reaches the target speed within the distance."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle.sim_ref import lib


def _probe(kind, a, b=0.0, c=0.0, d=0.0):
    f = lib().ref_probe_krauss
    f.restype = C.c_float
    f.argtypes = [C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float]
    return float(f(kind, a, b, c, d))


B, TAU = 10.0, 1.0          # decel of the vType (build_file.py:279), SUMO's default tau


def _distance_until_stop(v, b):
    """Euler braking from speed v: the vehicle moves with its NEW speed each second."""
    dist = 0.0
    while v > 0.0:
        v = max(0.0, v - b)
        dist += v
    return dist


@given(st.floats(0.0, 60.0))
def test_brake_gap_is_the_euler_braking_distance(v):
    assert abs(_probe(0, v, B) - _distance_until_stop(np.float32(v), B)) < 1e-2 + 1e-4 * v * v


@settings(max_examples=300)
@given(st.floats(0.0, 400.0), st.sampled_from([4.5, 10.0]))
def test_stop_speed_never_overshoots_the_gap(gap, b):
    v = _probe(1, gap, b, TAU)
    assert v >= 0.0
    # reaction time: the vehicle covers v * tau, then brakes
    travelled = v * TAU + _distance_until_stop(np.float32(v), b)
    assert travelled <= gap + 1e-2 + 1e-4 * gap


@given(st.floats(0.0, 300.0), st.floats(0.0, 50.0))
def test_stop_speed_is_monotone_in_the_gap(gap, extra):
    assert _probe(1, gap + extra, B, TAU) >= _probe(1, gap, B, TAU) - 1e-4


@settings(max_examples=300)
@given(st.floats(0.0, 200.0), st.floats(0.0, 30.0))
def test_follow_speed_is_safe_behind_a_braking_leader(gap, v_lead):
    """follower at follow_speed(gap, v_lead), leader braking with b from v_lead: the follower, braking after tau, stops
    before the leader's stopping point (net gap never negative)."""
    v = _probe(2, gap, v_lead, B, TAU)
    lead_travel = _distance_until_stop(np.float32(v_lead), B)
    own_travel = v * TAU + _distance_until_stop(np.float32(v), B)
    assert own_travel <= gap + lead_travel + 1e-2 + 1e-4 * (gap + lead_travel)
    # and a faster leader never forces a lower speed
    assert _probe(2, gap, v_lead + 5.0, B, TAU) >= v - 1e-4


@given(st.floats(0.0, 300.0), st.floats(1.0, 15.0))
def test_free_speed_reaches_the_target_within_the_distance(dist, target):
    v = _probe(3, dist, target, B)
    assert v >= target - 1e-4
    # braking from v towards `target` with b per second needs at most `dist`
    d, cur = 0.0, v
    while cur > target + 1e-3:
        cur = max(target, cur - B)
        d += cur
    assert d <= dist + target + 1e-2 + 1e-4 * dist

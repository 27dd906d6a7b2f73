"""ctypes wrapper of the CPU oracle (oracle/tsc_sim_ref.c).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from the
product package (deeprl_signal_control_b200 must not import this module).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libtsc_ref.so")
    src = os.path.join(_HERE, "tsc_sim_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.ref_create.restype = C.c_void_p
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class RefSim:
    """R independent replicas stepped sequentially (or with `threads` pthreads) on the CPU."""

    def __init__(self, net, params, n_replicas: int):
        self.net, self.params, self.R = net, params, int(n_replicas)
        self._cnet, self._ccfg = net.as_c(), params.as_c()
        self.h = C.c_void_p(lib().ref_create(C.byref(self._cnet), C.byref(self._ccfg), self.R))

    def __del__(self):
        try:
            lib().ref_destroy(self.h)
        except Exception:
            pass

    def reset(self, seeds):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint64)
        assert seeds.shape == (self.R,)
        lib().ref_reset(self.h, _p(seeds, C.c_uint64))

    def set_train_mode(self, m: bool):
        lib().ref_set_train_mode(self.h, int(m))

    def observe(self, fp=None):
        n = self.net
        obs = np.zeros((self.R, n.n_obs), np.float32)
        fp = None if fp is None else np.ascontiguousarray(fp, np.float32)
        lib().ref_observe(self.h, _p(fp, C.c_float), _p(obs, C.c_float))
        return obs

    def step(self, action, fp=None, threads: int = 1):
        n = self.net
        action = np.ascontiguousarray(action, np.int32).reshape(self.R, n.n_nodes)
        fp = None if fp is None else np.ascontiguousarray(fp, np.float32)
        obs = np.zeros((self.R, n.n_obs), np.float32)
        reward = np.zeros((self.R, n.n_nodes), np.float32)
        greward = np.zeros(self.R, np.float32)
        done = np.zeros(self.R, np.uint8)
        lib().ref_step_mt(self.h, _p(action, C.c_int32), _p(fp, C.c_float), _p(obs, C.c_float),
                          _p(reward, C.c_float), _p(greward, C.c_float), _p(done, C.c_uint8),
                          C.c_int32(threads))
        return obs, reward, greward, done

    def run(self, actions, n_steps: int, fp=None, threads: int = 1):
        """n_steps control steps in ONE call (actions [n_act, R, n_nodes] cycled): each thread walks its replica range
        through all steps, so the thread pool is created once.  Returns the last step's (obs, reward, greward, done)."""
        n = self.net
        actions = np.ascontiguousarray(actions, np.int32).reshape(-1, self.R, n.n_nodes)
        fp = None if fp is None else np.ascontiguousarray(fp, np.float32)
        obs = np.zeros((self.R, n.n_obs), np.float32)
        reward = np.zeros((self.R, n.n_nodes), np.float32)
        greward = np.zeros(self.R, np.float32)
        done = np.zeros(self.R, np.uint8)
        lib().ref_run_mt(self.h, _p(actions, C.c_int32), C.c_int32(len(actions)), C.c_int32(n_steps), _p(fp, C.c_float),
                         _p(obs, C.c_float), _p(reward, C.c_float), _p(greward, C.c_float), _p(done, C.c_uint8),
                         C.c_int32(threads))
        return obs, reward, greward, done

    def counts(self):
        n = self.net
        veh = np.zeros((self.R, n.n_det), np.int32)
        halt = np.zeros_like(veh)
        wait = np.zeros_like(veh)
        phase = np.zeros((self.R, n.n_nodes), np.int32)
        lib().ref_get_counts(self.h, _p(veh, C.c_int32), _p(halt, C.c_int32), _p(wait, C.c_int32),
                             _p(phase, C.c_int32))
        return veh, halt, wait, phase

    def dump_state(self, replica: int = 0):
        n = self.net
        cnt = np.zeros(n.n_lanes, np.int32)
        veh = np.zeros((n.n_slots, 3), np.uint32)
        nv = C.c_int32(0)
        lib().ref_dump_state(self.h, C.c_int32(replica), _p(cnt, C.c_int32), _p(veh, C.c_uint32),
                             C.byref(nv))
        return cnt, veh[:nv.value].copy()

    def misc(self, replica: int = 0):
        out = np.zeros(5, np.int32)
        lib().ref_get_misc(self.h, C.c_int32(replica), _p(out, C.c_int32))
        return dict(cur_sec=int(out[0]), departed=int(out[1]), arrived=int(out[2]),
                    backlog=int(out[3]), live=int(out[4]))

    def backlog(self, replica: int = 0):
        out = np.zeros(self.net.n_src, np.int32)
        lib().ref_get_backlog(self.h, C.c_int32(replica), _p(out, C.c_int32))
        return out

    # ---- evaluation / recording path (envs/env.py:409-437, 498-542) --------------------------
    def set_record(self, on: bool = True):
        lib().ref_set_record(self.h, int(on))

    def traffic_stats(self):
        out = np.zeros((self.R, 8), np.float32)
        lib().ref_traffic_stats(self.h, _p(out, C.c_float))
        return out

    def step_record(self, action, fp=None):
        """step() one simulated second at a time; also returns the per-second traffic statistics
        [R, control_interval_sec, 8]."""
        n = self.net
        action = np.ascontiguousarray(action, np.int32).reshape(self.R, n.n_nodes)
        fp = None if fp is None else np.ascontiguousarray(fp, np.float32)
        obs = np.zeros((self.R, n.n_obs), np.float32)
        reward = np.zeros((self.R, n.n_nodes), np.float32)
        greward = np.zeros(self.R, np.float32)
        done = np.zeros(self.R, np.uint8)
        stats = np.zeros((self.R, self.params.control_interval_sec, 8), np.float32)
        lib().ref_step_record(self.h, _p(action, C.c_int32), _p(fp, C.c_float), _p(obs, C.c_float),
                              _p(reward, C.c_float), _p(greward, C.c_float), _p(done, C.c_uint8),
                              _p(stats, C.c_float))
        return obs, reward, greward, done, stats

    def trips(self, replica: int = 0):
        """tripinfo rows of one replica: int array [n, 5] = depart_sec, arrival_sec, route, wait_sec, wait_count."""
        rows = np.zeros((8192, 2), np.uint32)
        nr = C.c_int32(0)
        lib().ref_get_trips(self.h, C.c_int32(replica), _p(rows, C.c_uint32), C.c_int32(len(rows)), C.byref(nr))
        return decode_trips(rows[:nr.value])


def decode_trips(rows: np.ndarray) -> np.ndarray:
    w0, w1 = rows[:, 0].astype(np.int64), rows[:, 1].astype(np.int64)
    return np.stack([w0 & 4095, (w0 >> 12) & 4095, w0 >> 24, w1 & 65535, w1 >> 16], axis=1)

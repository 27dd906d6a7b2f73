"""Replay the action traces the reference RECORDED on Monaco (real_net_experimental_data/eva_data/*_control.csv: the
28-element action vector of every control step of 10 evaluation episodes per agent, produced by the reference's trained
agents on SUMO) through OUR simulator (CPU oracle = the CUDA kernel bit for bit) and compare the aggregate traffic it
produces with what SUMO produced under the same signal plans (`*_traffic.csv`, `*_trip.csv`).

This is a distribution-level check of the restated dynamics: the vehicle model is NOT pinned against SUMO,
so the numbers quantify the gap instead of asserting equality.  Needs /root/reference; writes
tests/golden/monaco_replay_summary.json.      Run: python tests/golden/replay_monaco_eval_traces.py
"""
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/real_net_experimental_data/eva_data"

from deeprl_signal_control_b200.net.real_net import real_net_tables    # noqa: E402
from oracle.sim_ref import RefSim                                       # noqa: E402
from tests.test_real_net_cpu import real_params                         # noqa: E402


TAU = 1.0          # set by __main__: 1.0 = the current reference (README.md:63 removed tau="0.5"), 0.5 = the paper-era vType


def _params():
    par = real_params("greedy")
    par.tau = TAU
    return par


def replay(agent, n_episodes=10):
    ctrl = pd.read_csv(os.path.join(REF, "real_net_%s_control.csv" % agent))
    traffic = pd.read_csv(os.path.join(REF, "real_net_%s_traffic.csv" % agent))
    trips = pd.read_csv(os.path.join(REF, "real_net_%s_trip.csv" % agent))
    net, par = real_net_tables("greedy"), _params()
    eps = sorted(ctrl.episode.unique())[:n_episodes]
    R = len(eps)
    acts = np.stack([np.array([[int(x) for x in s.split(",")] for s in ctrl[ctrl.episode == e].action], np.int32)
                     for e in eps], axis=1)                              # [720, R, 28]
    assert acts.shape[2] == net.n_nodes and (acts.max(axis=(0, 1)) < np.asarray(net.n_a_ls)).all()
    sim = RefSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64) + np.uint64(10000))
    sim.set_train_mode(False)
    sim.set_record(True)
    stats, rewards = [], []
    for t in range(acts.shape[0]):
        _, rew, _, _, st = sim.step_record(acts[t])
        stats.append(st); rewards.append(rew.sum(1))
    stats = np.concatenate(stats, axis=1)                                # [R, 3600, 8]
    ours_trips = np.concatenate([sim.trips(r) for r in range(R)])
    dur = ours_trips[:, 1] - ours_trips[:, 0]
    ref_t = traffic[traffic.episode.isin(eps)]
    ref_trips = trips[trips.episode.isin(eps)]
    return {
        "agent": agent, "episodes": R,
        "ours": {"avg_queue": float(stats[..., 5].mean()), "avg_speed_mps": float(stats[..., 4].mean()),
                 "avg_wait_sec": float(stats[..., 3].mean()), "peak_cars": float(stats[..., 0].max(1).mean()),
                 "trips_per_episode": len(ours_trips) / R, "mean_trip_duration_sec": float(dur.mean()),
                 "mean_trip_wait_sec": float(ours_trips[:, 3].mean()),
                 "departed_per_episode": float(stats[:, -1, 1].mean()),
                 "sum_local_reward_per_step": float(np.mean(rewards))},
        "sumo_recorded": {"avg_queue": float(ref_t.avg_queue.mean()), "avg_speed_mps": float(ref_t.avg_speed_mps.mean()),
                          "avg_wait_sec": float(ref_t.avg_wait_sec.mean()),
                          "peak_cars": float(ref_t.groupby("episode").number_total_car.max().mean()),
                          "trips_per_episode": len(ref_trips) / R,
                          "mean_trip_duration_sec": float(ref_trips.duration_sec.mean()),
                          "mean_trip_wait_sec": float(ref_trips.wait_sec.mean()),
                          "departed_per_episode": float(ref_t.groupby("episode").number_departed_car.sum().mean()),
                          "sum_local_reward_per_step": float(ctrl[ctrl.episode.isin(eps)].reward.mean())},
    }


def closed_loop_greedy(n_episodes=10):
    """OUR greedy controller driving OUR simulator (closed loop), against SUMO's recorded greedy episodes: the fair
    comparison of the two traffic models under the same control LAW (the open-loop replays above apply plans that were
    reactive to SUMO's traffic, not ours)."""
    from deeprl_signal_control_b200.envs.env import Node
    from deeprl_signal_control_b200.envs.real_net_env import RealNetController
    traffic = pd.read_csv(os.path.join(REF, "real_net_greedy_traffic.csv"))
    trips = pd.read_csv(os.path.join(REF, "real_net_greedy_trip.csv"))
    ctrl_csv = pd.read_csv(os.path.join(REF, "real_net_greedy_control.csv"))
    net, par = real_net_tables("greedy"), _params()
    nodes = {}
    for name in net.node_names:
        nd = Node(name)
        nd.lanes_in, nd.ilds_in = net.lanes_in[name], net.ilds_in[name]
        nodes[name] = nd
    ctrl = RealNetController(net.node_names, nodes)
    R = n_episodes
    sim = RefSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64) + np.uint64(10000))
    sim.set_train_mode(False)
    sim.set_record(True)
    off = net.node_obs_off
    obs = sim.observe()
    stats, rewards = [], []
    for t in range(720):
        act = np.array([ctrl.forward([obs[r][off[i]:off[i + 1]] for i in range(net.n_nodes)]) for r in range(R)], np.int32)
        obs, rew, _, _, st = sim.step_record(act)
        stats.append(st); rewards.append(rew.sum(1))
    stats = np.concatenate(stats, axis=1)
    ours_trips = np.concatenate([sim.trips(r) for r in range(R)])
    dur = ours_trips[:, 1] - ours_trips[:, 0]
    return {"agent": "greedy (closed loop: our controller on our simulator)", "episodes": R,
            "ours": {"avg_queue": float(stats[..., 5].mean()), "avg_speed_mps": float(stats[..., 4].mean()),
                     "avg_wait_sec": float(stats[..., 3].mean()), "peak_cars": float(stats[..., 0].max(1).mean()),
                     "trips_per_episode": len(ours_trips) / R, "mean_trip_duration_sec": float(dur.mean()),
                     "mean_trip_wait_sec": float(ours_trips[:, 3].mean()),
                     "departed_per_episode": float(stats[:, -1, 1].mean()),
                     "sum_local_reward_per_step": float(np.mean(rewards))},
            "sumo_recorded": {"avg_queue": float(traffic.avg_queue.mean()), "avg_speed_mps": float(traffic.avg_speed_mps.mean()),
                              "avg_wait_sec": float(traffic.avg_wait_sec.mean()),
                              "peak_cars": float(traffic.groupby("episode").number_total_car.max().mean()),
                              "trips_per_episode": len(trips) / traffic.episode.nunique(),
                              "mean_trip_duration_sec": float(trips.duration_sec.mean()),
                              "mean_trip_wait_sec": float(trips.wait_sec.mean()),
                              "departed_per_episode": float(traffic.groupby("episode").number_departed_car.sum().mean()),
                              "sum_local_reward_per_step": float(ctrl_csv.reward.mean())}}


if __name__ == "__main__":
    result = {"what": __doc__.split("\n\n")[0].replace("\n", " "), "by_tau": {}}
    for tau in (1.0, 0.5):
        TAU = tau
        out = [closed_loop_greedy()] + [replay(a) for a in ("greedy", "ma2c", "ia2c")]
        result["by_tau"]["%.1f" % tau] = out
        print("=== tau = %.1f ===" % tau)
        for r in out:
            print(r["agent"])
            for k in r["ours"]:
                print("   %-28s ours %9.3f   sumo %9.3f" % (k, r["ours"][k], r["sumo_recorded"][k]))
    json.dump(result, open(os.path.join(HERE, "monaco_replay_summary.json"), "w"), indent=1)

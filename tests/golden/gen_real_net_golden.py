"""Golden vectors for the Monaco `real_net` scenario from the REFERENCE's own Python
(envs/env.py + envs/real_net_env.py executed over a fake TraCI connection backed by our oracle; see
gen_env_golden.py for the stubbing).  Pins: agent order, n_s_ls/n_a_ls/n_w_ls/n_f_ls, ilds_in (i.e. our
net-file ingest of getControlledLanes), the 'queue' objective with min(10, halting) per lane
(envs/env.py:333), reward normalisation by REALNET_REWARD_NORM (envs/env.py:599-601,625-629),
the wave-only state and RealNetController.greedy.   Run: python tests/golden/gen_real_net_golden.py
"""
import configparser
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_env_golden as G          # noqa: E402  (installs the stubs, sys.path entries)

from envs.real_net_env import RealNetEnv, RealNetController      # noqa: E402  (reference)
from deeprl_signal_control_b200.net.real_net import build_real_net   # noqa: E402
from deeprl_signal_control_b200.net.tables import EnvParams          # noqa: E402
from oracle.sim_ref import RefSim                                    # noqa: E402

NET_FILE = os.path.join(G.REF, "real_net", "data", "in", "most.net.xml")


def real_params(agent):
    return EnvParams(agent=agent, objective="queue", norm_wave=5.0, norm_wait=100.0, clip_wave=2.0, clip_wait=2.0,
                     coef_wait=0.0, coop_gamma=0.9, teleport_sec=300, real_net_norm=True, use_wait=False,
                     det_len=-1.0, halt_speed=0.1, queue_cap=10)


class FakeTraciReal(G.FakeTraci):
    """Whole-lane `lane.*` queries (envs/env.py:333,341,377,388)."""

    def __init__(self, net, params, seed):
        super().__init__(net, params, seed)
        s = self
        self.trafficlight.getIDList = lambda: list(net.node_names)
        self.lane = G._NS()
        self.lane.getLastStepVehicleNumber = lambda ild: len(s.det[ild])
        self.lane.getLastStepHaltingNumber = lambda ild: sum(1 for v in s.det[ild] if v[1] < 0.1)
        self.lane.getLastStepVehicleIDs = lambda ild: ["%s#%d" % (ild, k) for k in range(len(s.det[ild]))]

    def _refresh(self):
        cnt, veh = self.ref.dump_state(0)
        self.det, k = {}, 0
        for l, c in enumerate(cnt):
            rows = veh[k:k + c]; k += c
            pos = rows[:, 0].copy().view(np.float32); spd = rows[:, 1].copy().view(np.float32)
            wait = (rows[:, 2] & 1023).astype(np.int64)
            self.det[self.net.lane_names[l]] = [(float(p), float(s_), int(w)) for p, s_, w in zip(pos, spd, wait)]


class GoldenRealEnv(RealNetEnv):
    def _init_sim(self, seed, gui=False):
        self.sim = FakeTraciReal(self._tables, self._params, seed)


def run(agent, train_mode, n_steps):
    cp = configparser.ConfigParser()
    cp.read(os.path.join(G.REF, "config", "config_ma2c_real.ini"))
    cfg = cp["ENV_CONFIG"]
    cfg["agent"] = agent
    net = build_real_net(NET_FILE, flow_rate=cfg.getint("flow_rate"), agent=agent, coop_gamma=cfg.getfloat("coop_gamma"))
    GoldenRealEnv._tables, GoldenRealEnv._params = net, real_params(agent)
    env = GoldenRealEnv(cfg)
    env.train_mode = train_mode
    rng = np.random.default_rng(11)
    ob = env.reset()
    fake = env.sim
    ctrl = RealNetController(env.node_names, env.nodes)
    out = dict(actions=[], fps=[], obs=[np.concatenate(ob)], reward=[], greward=[], done=[], greedy=[])
    na = np.array(env.n_a_ls)
    for t in range(n_steps):
        if agent == "greedy":
            act = np.array(ctrl.forward(ob), dtype=np.int32)
            out["greedy"].append(act.copy())
            if t % 4 == 3:
                act = (rng.integers(0, 1 << 20, len(na)) % na).astype(np.int32)
        else:
            act = (rng.integers(0, 1 << 20, len(na)) % na).astype(np.int32)
        fp_full = np.zeros((len(na), net.max_na), np.float32)
        if agent == "ma2c":
            pol = []
            for i, n_a in enumerate(na):
                p = rng.dirichlet(np.ones(n_a)).astype(np.float32)
                pol.append(p); fp_full[i, :n_a] = p
            env.update_fingerprint(pol)
        fake.pending_action = act.reshape(1, -1)
        fake.pending_fp = fp_full[None] if agent == "ma2c" else None
        ob, reward, done, greward = env.step(list(act))
        out["actions"].append(act); out["fps"].append(fp_full)
        out["obs"].append(np.concatenate(ob))
        out["reward"].append(np.asarray(reward, dtype=np.float64) * np.ones(len(na)))
        out["greward"].append(float(greward)); out["done"].append(bool(done))
    meta = dict(node_names=env.node_names, n_s_ls=[int(x) for x in env.n_s_ls], n_a_ls=[int(x) for x in env.n_a_ls],
                n_w_ls=[int(x) for x in env.n_w_ls], n_f_ls=[int(x) for x in env.n_f_ls], T=float(env.T),
                seed0=fake.seed, ilds_in={k: v.ilds_in for k, v in env.nodes.items()},
                neighbor={k: v.neighbor for k, v in env.nodes.items()})
    return out, meta


if __name__ == "__main__":
    for agent, train, n in [("ma2c", True, 140), ("ia2c", True, 80), ("greedy", False, 120)]:
        out, meta = run(agent, train, n)
        tag = "%s_%s" % (agent, "train" if train else "test")
        np.savez_compressed(os.path.join(HERE, "real_%s.npz" % tag),
                            actions=np.array(out["actions"], np.int32), fps=np.array(out["fps"], np.float32),
                            obs=np.array(out["obs"], np.float64), reward=np.array(out["reward"], np.float64),
                            greward=np.array(out["greward"], np.float64), done=np.array(out["done"]),
                            greedy=np.array(out["greedy"], np.int32) if out["greedy"] else np.zeros(0, np.int32),
                            meta=json.dumps(meta))
        print(tag, "mean greward", np.mean(out["greward"]), "obs dim", len(out["obs"][0]), "min greward", np.min(out["greward"]))

"""Generate golden vectors by executing the REFERENCE's own Python (envs/env.py,
envs/large_grid_env.py, large_grid/data/build_file.py, agents/utils.py) in this container.

SUMO/TraCI/TensorFlow are not installed, so:
  * `traci`, `sumolib`, `tensorflow`, `matplotlib`, `seaborn` are stubbed as empty modules;
  * the TraCI connection object (`env.sim`) is replaced by `FakeTraci`, which serves the
    detector / vehicle queries from OUR CPU oracle's vehicle state and records every
    `setRedYellowGreenState` string the reference emits.
The reference code that runs unmodified: TrafficSimulator.__init__/_init_nodes/_init_state_space/
reset/step/_get_node_phase/_set_phase/_measure_state_step/_measure_reward_step/_get_state/
reward shaping/update_fingerprint, LargeGridPhase, LargeGridController.greedy, and the XML
generators of build_file.py (edges, connections, detectors, flows), OnPolicyBuffer, Scheduler.

Outputs (committed): tests/golden/env_<agent>.npz, grid_spec.json, buffers.npz.
Run:  python tests/golden/gen_env_golden.py      (needs /root/reference; NOT needed at test time)
"""
import configparser
import json
import os
import re
import sys
import types
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

# ---- stubs for modules that are absent here -------------------------------------------------
for name in ["traci", "sumolib", "seaborn", "matplotlib", "matplotlib.pyplot", "tensorflow"]:
    m = types.ModuleType(name)
    sys.modules[name] = m
sys.modules["sumolib"].checkBinary = lambda x: x
sys.modules["seaborn"].set_color_codes = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["xml.etree.cElementTree"] = ET          # removed in Python 3.9


class _Any:
    """tensorflow stand-in: any attribute / call yields another stand-in (only default arguments
    such as `act=tf.nn.relu` are evaluated when agents/utils.py is imported)."""

    def __getattr__(self, k):
        return _Any()

    def __call__(self, *a, **k):
        return _Any()


sys.modules["tensorflow"] = _Any()

from envs.large_grid_env import LargeGridEnv, LargeGridController   # noqa: E402  (reference)
import large_grid.data.build_file as ref_build                      # noqa: E402  (reference)

from deeprl_signal_control_b200.net.large_grid import build_large_grid  # noqa: E402
from deeprl_signal_control_b200.net.tables import EnvParams              # noqa: E402
from oracle.sim_ref import RefSim                                        # noqa: E402


class _NS:
    pass


class FakeTraci:
    """Serves the TraCI calls of envs/env.py from the oracle's state."""

    def __init__(self, net, params, seed):
        self.net = net
        self.seed = int(seed)   # train: config seed (+episode); test: test_seeds[k] (envs/env.py:547-550)
        self.ref = RefSim(net, params, 1)
        self.ref.reset([seed])
        self.ref.set_train_mode(False)
        self.lane_idx = {n: i for i, n in enumerate(net.lane_names)}
        self.pending_action = None
        self.pending_fp = None
        self.n_sim = 0
        self.phase_log = []      # (node, state string) in call order
        self._refresh()
        s = self
        self.trafficlight = _NS()
        self.trafficlight.getIDList = lambda: ["nt%d" % i for i in range(1, 26)]
        self.trafficlight.getControlledLanes = lambda node: list(net.lanes_in[node])
        self.trafficlight.setRedYellowGreenState = lambda node, st: s.phase_log.append((node, st))
        self.trafficlight.setPhaseDuration = lambda node, d: None
        self.lanearea = _NS()
        self.lanearea.getLastStepVehicleNumber = lambda ild: len(s.det[ild])
        self.lanearea.getLastStepHaltingNumber = lambda ild: sum(1 for v in s.det[ild] if v[1] < 1.39)
        self.lanearea.getLastStepVehicleIDs = lambda ild: ["%s#%d" % (ild, k) for k in range(len(s.det[ild]))]
        self.vehicle = _NS()
        self.vehicle.getLanePosition = lambda vid: s._veh(vid)[0]
        self.vehicle.getWaitingTime = lambda vid: s._veh(vid)[2]

    def _veh(self, vid):
        ild, k = vid.rsplit("#", 1)
        return self.det[ild][int(k)]

    def _refresh(self):
        cnt, veh = self.ref.dump_state(0)
        self.det = {}
        k = 0
        for l, c in enumerate(cnt):
            L = float(self.net.lane_len[l])
            rows = veh[k:k + c]
            k += c
            pos = rows[:, 0].copy().view(np.float32)
            spd = rows[:, 1].copy().view(np.float32)
            wait = (rows[:, 2] & 1023).astype(np.int64)
            # E2 detector pos=-50 endPos=-1 (build_file.py:445): vehicles whose front is inside
            self.det[self.net.lane_names[l]] = [(float(p), float(s_), int(w)) for p, s_, w in zip(pos, spd, wait)
                                                if p > L - 50.0]

    def simulationStep(self):
        self.n_sim += 1
        if self.n_sim % 5 == 0:          # control_interval_sec simulated seconds have elapsed
            self.ref.step(self.pending_action, self.pending_fp)
            self._refresh()

    def close(self):
        pass


class GoldenEnv(LargeGridEnv):
    """Reference env with only the SUMO process launch replaced."""

    def _init_sim(self, seed, gui=False):
        self.sim = FakeTraci(self._tables, self._params, seed)


def run(agent, train_mode, n_steps, seed_offset=0):
    cfgp = configparser.ConfigParser()
    cfgp.read(os.path.join(REF, "config", "config_ma2c_large.ini"))
    cfg = cfgp["ENV_CONFIG"]
    cfg["agent"] = agent
    net = build_large_grid(agent=agent, coop_gamma=cfg.getfloat("coop_gamma"))
    params = EnvParams(agent=agent)
    GoldenEnv._tables, GoldenEnv._params = net, params
    env = GoldenEnv(cfg)
    env.train_mode = train_mode
    env.seed += seed_offset
    rng = np.random.default_rng(7)
    ob = env.reset()
    fake = env.sim
    out = dict(actions=[], fps=[], obs=[np.concatenate(ob)], reward=[], greward=[], done=[],
               yellow=[], green=[], greedy=[])
    ctrl = LargeGridController(env.node_names)
    for t in range(n_steps):
        if agent == "greedy":
            act = np.array(ctrl.forward(ob), dtype=np.int32)
            out["greedy"].append(act.copy())
            if t % 3 == 2:          # perturb so that yellow/no-yellow cases both appear
                act = rng.integers(0, 5, 25).astype(np.int32)
        else:
            act = rng.integers(0, 5, 25).astype(np.int32)
            if t % 4 == 0:
                act = out["actions"][-1].copy() if out["actions"] else act   # repeated action: no yellow
        fp_full = np.zeros((25, 5), np.float32)
        if agent == "ma2c":
            pol = rng.dirichlet(np.ones(5), size=25).astype(np.float32)
            env.update_fingerprint(list(pol))
            fp_full[:] = pol
        fake.pending_action = act.reshape(1, 25)
        fake.pending_fp = fp_full.reshape(1, 25, 5) if agent == "ma2c" else None
        n0 = len(fake.phase_log)
        ob, reward, done, greward = env.step(list(act))
        log = fake.phase_log[n0:]
        assert len(log) == 50
        out["yellow"].append([s for _, s in log[:25]])
        out["green"].append([s for _, s in log[25:]])
        out["actions"].append(act); out["fps"].append(fp_full)
        out["obs"].append(np.concatenate(ob) if agent != "greedy" else np.concatenate(ob))
        out["reward"].append(np.asarray(reward, dtype=np.float64) * np.ones(25))
        out["greward"].append(float(greward)); out["done"].append(bool(done))
    meta = dict(node_names=env.node_names, n_s_ls=[int(x) for x in env.n_s_ls],
                n_a_ls=[int(x) for x in env.n_a_ls], n_w_ls=[int(x) for x in env.n_w_ls],
                n_f_ls=[int(x) for x in env.n_f_ls], T=float(env.T), seed0=fake.seed,
                ilds_in={k: v.ilds_in for k, v in env.nodes.items()},
                neighbor={k: v.neighbor for k, v in env.nodes.items()})
    return out, meta


def grid_spec():
    """Parse the XML the reference generator emits into plain lists (edges, connections,
    detectors, flows) — the structural golden for net/large_grid.py."""
    edges = re.findall(r'<edge id="(\S+)" from="(\S+)" to="(\S+)" type="(\S+)"/>',
                       ref_build.output_edges('  <edge id="%s" from="%s" to="%s" type="%s"/>\n'))
    cons = re.findall(r'from="(\S+)" to="(\S+)" fromLane="(\d)" toLane="(\d)"',
                      ref_build.output_connections('  <connection from="%s" to="%s" fromLane="%d" toLane="%d"/>\n'))
    ilds = re.findall(r'lane="(\S+)"', ref_build.output_ild('  <ild id="%s_%d" lane="%s_%d"/>\n'))
    flows = re.findall(r'<flow id="(\S+)" departPos="random_free" from="(\S+)" to="(\S+)" begin="(\d+)" '
                       r'end="(\d+)" vehsPerHour="(\d+)"', ref_build.output_flows(1100, 925, 0, seed=12))
    nodes = re.findall(r'<node id="(\S+)" x="(\S+)" y="(\S+)" type="(\S+)"/>',
                       ref_build.output_nodes('  <node id="%s" x="%.2f" y="%.2f" type="%s"/>\n'))
    # initial fleet of init_density = 0.2 (init_routes, large_grid/data/build_file.py:223-266), seeds 12 and 31
    init = {}
    for seed in (12, 31):
        xml = ref_build.output_flows(1100, 925, 0.2, seed=seed)
        init[str(seed)] = re.findall(r'<flow id="i_(\d+)" departPos="random_free" from="(\S+)" to="(\S+)" begin="0" end="1" '
                                     r'departLane="(\d)" departSpeed="0" number="(\d+)"', xml)
    return dict(edges=edges, connections=cons, ilds=ilds, flows=flows, nodes=nodes, init_fleet=init)


def buffers_golden():
    """OnPolicyBuffer / Scheduler of reference agents/utils.py (pure numpy once TF is stubbed)."""
    from agents.utils import OnPolicyBuffer, Scheduler
    rng = np.random.default_rng(3)
    out = {}
    buf = OnPolicyBuffer(0.99)
    batches = []
    done_prev = True
    buf.reset(done_prev)
    for b in range(3):
        T = 12
        rs = rng.normal(size=T); vs = rng.normal(size=T)
        dones = rng.random(T) < 0.15
        for t in range(T):
            buf.add_transition(np.zeros(3, np.float32), int(rng.integers(0, 5)), rs[t], vs[t], bool(dones[t]))
        R = float(rng.normal())
        obs, acts, d, Rs, Advs = buf.sample_transition(R)
        batches.append(dict(rs=rs, vs=vs, dones=dones, R=R, pre_dones=np.array(d, dtype=bool), Rs=Rs, Advs=Advs))
    for k in batches[0]:
        out["buf_" + k] = np.stack([np.asarray(bt[k]) for bt in batches])
    s = Scheduler(1.0, 0.01, 1000.0, decay="linear")
    out["sched_linear"] = np.array([s.get(120) for _ in range(12)])
    s = Scheduler(5e-4, decay="constant")
    out["sched_const"] = np.array([s.get(120) for _ in range(3)])
    return out


if __name__ == "__main__":
    for agent, train, n in [("ma2c", True, 160), ("ia2c", True, 100), ("greedy", False, 100), ("ma2c", False, 60)]:
        out, meta = run(agent, train, n)
        tag = "%s_%s" % (agent, "train" if train else "test")
        np.savez_compressed(os.path.join(HERE, "env_%s.npz" % tag),
                            actions=np.array(out["actions"], np.int32), fps=np.array(out["fps"], np.float32),
                            obs=np.array(out["obs"], np.float64), reward=np.array(out["reward"], np.float64),
                            greward=np.array(out["greward"], np.float64), done=np.array(out["done"]),
                            yellow=np.array(out["yellow"]), green=np.array(out["green"]),
                            greedy=np.array(out["greedy"], np.int32) if out["greedy"] else np.zeros(0, np.int32),
                            meta=json.dumps(meta))
        print(tag, "mean greward", np.mean(out["greward"]), "obs dim", len(out["obs"][0]))
    with open(os.path.join(HERE, "grid_spec.json"), "w") as f:
        json.dump(grid_spec(), f)
    np.savez_compressed(os.path.join(HERE, "buffers.npz"), **buffers_golden())
    print("done")

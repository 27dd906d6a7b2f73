"""Golden vectors for the 6-intersection `small_grid` scenario from the REFERENCE's own Python
(envs/env.py + envs/small_grid_env.py executed over a fake TraCI connection backed by our oracle; see
gen_env_golden.py for the stubbing).  The reference supports this scenario for the greedy / a2c agents only (its
neighbour map names the non-TLS junction `npc`), so the goldens are for `greedy`.  Pins: agent order and per-node
phase counts (3 for nt1, 2 elsewhere), ilds_in from our link order, the un-normalised wave state (norm 1.0, clip
1000), the hybrid reward with E2 halting counts, yellow strings for 2- and 3-phase nodes, SmallGridController.greedy
with STATE_PHASE_MAP.     Run: python tests/golden/gen_small_grid_golden.py
"""
import configparser
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_env_golden as G          # noqa: E402  (installs the stubs, sys.path entries)

sys.modules["matplotlib"].use = lambda *a, **k: None      # envs/small_grid_env.py:10

from envs.small_grid_env import SmallGridEnv, SmallGridController      # noqa: E402  (reference)
from deeprl_signal_control_b200.net.small_grid import build_small_grid  # noqa: E402
from deeprl_signal_control_b200.net.tables import EnvParams             # noqa: E402


def small_params(agent, cfg):
    return EnvParams(agent=agent, objective=cfg.get("objective"), norm_wave=cfg.getfloat("norm_wave"),
                     norm_wait=cfg.getfloat("norm_wait"), clip_wave=cfg.getfloat("clip_wave"),
                     clip_wait=cfg.getfloat("clip_wait"), coef_wait=cfg.getfloat("coef_wait"),
                     coop_gamma=cfg.getfloat("coop_gamma"))


class FakeTraciSmall(G.FakeTraci):
    def __init__(self, net, params, seed):
        super().__init__(net, params, seed)
        self.trafficlight.getIDList = lambda: list(net.node_names)


class GoldenSmallEnv(SmallGridEnv):
    def _init_sim(self, seed, gui=False):
        self.sim = FakeTraciSmall(self._tables, self._params, seed)


def run(n_steps):
    cp = configparser.ConfigParser()
    cp.read(os.path.join(G.REF, "config", "config_test_small.ini"))
    cfg = cp["ENV_CONFIG"]
    cfg["scenario"] = "small_grid"      # the shipped file says large_grid; main.py:52 dispatches on 'small_grid'
    agent = cfg.get("agent")
    net = build_small_grid(cfg.getint("num_extra_car_per_hour"), agent=agent, coop_gamma=cfg.getfloat("coop_gamma"))
    GoldenSmallEnv._tables, GoldenSmallEnv._params = net, small_params(agent, cfg)
    env = GoldenSmallEnv(cfg)
    env.train_mode = False
    rng = np.random.default_rng(5)
    ob = env.reset()
    fake = env.sim
    ctrl = SmallGridController(env.node_names)
    na = np.array(env.n_a_ls)
    out = dict(actions=[], obs=[np.concatenate(ob)], reward=[], greward=[], done=[], greedy=[], yellow=[], green=[])
    for t in range(n_steps):
        act = np.array(ctrl.forward(ob), dtype=np.int32)
        out["greedy"].append(act.copy())
        if t % 3 == 2:
            act = (rng.integers(0, 1 << 20, len(na)) % na).astype(np.int32)
        fake.pending_action = act.reshape(1, -1)
        fake.pending_fp = None
        n0 = len(fake.phase_log)
        ob, reward, done, greward = env.step(list(act))
        log = fake.phase_log[n0:]
        assert len(log) == 2 * len(na)
        out["yellow"].append([s for _, s in log[:len(na)]]); out["green"].append([s for _, s in log[len(na):]])
        out["actions"].append(act); out["obs"].append(np.concatenate(ob))
        out["reward"].append(np.asarray(reward, dtype=np.float64) * np.ones(len(na)))
        out["greward"].append(float(greward)); out["done"].append(bool(done))
    meta = dict(node_names=env.node_names, n_s_ls=[int(x) for x in env.n_s_ls], n_a_ls=[int(x) for x in env.n_a_ls],
                n_w_ls=[int(x) for x in env.n_w_ls], n_f_ls=[int(x) for x in env.n_f_ls], T=float(env.T),
                seed0=fake.seed, ilds_in={k: v.ilds_in for k, v in env.nodes.items()},
                yellow=out["yellow"], green=out["green"],
                cfg={k: cfg.get(k) for k in cfg})
    return out, meta


if __name__ == "__main__":
    out, meta = run(200)
    np.savez_compressed(os.path.join(HERE, "small_greedy_test.npz"),
                        actions=np.array(out["actions"], np.int32), obs=np.array(out["obs"], np.float64),
                        reward=np.array(out["reward"], np.float64), greward=np.array(out["greward"], np.float64),
                        done=np.array(out["done"]), greedy=np.array(out["greedy"], np.int32), meta=json.dumps(meta))
    print("small greedy: mean greward", np.mean(out["greward"]), "min", np.min(out["greward"]), "obs dim", len(out["obs"][0]))

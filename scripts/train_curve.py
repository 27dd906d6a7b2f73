"""Short MA2C / IA2C training run (device-resident loop) to check that the learner learns:
mean episode reward (mean over the 720 control steps of the global reward, utils.py:296-305) per episode,
averaged over replicas.  Reference points of OUR restated simulator on the 5x5 grid (seed 12): random
policy about -315, greedy controller about -140.

  python scripts/train_curve.py R EPISODES AGENT [grid|real] [lstm|fc]
`real` = Monaco (config/config_ma2c_real.ini: n_step 40, reward_norm 1.0, queue objective)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from deeprl_signal_control_b200.agents.layout import PolicyLayout
from deeprl_signal_control_b200.agents.learner import BatchedA2C
from deeprl_signal_control_b200.agents.trainer import BatchedTrainer
from deeprl_signal_control_b200.net.large_grid import build_large_grid
from deeprl_signal_control_b200.net.tables import EnvParams
from deeprl_signal_control_b200.sim import BatchedSim

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
episodes = int(sys.argv[2]) if len(sys.argv) > 2 else 24
agent = sys.argv[3] if len(sys.argv) > 3 else "ma2c"
scenario = sys.argv[4] if len(sys.argv) > 4 else "grid"
policy = sys.argv[5] if len(sys.argv) > 5 else "lstm"
if scenario == "real":
    from deeprl_signal_control_b200.net.real_net import real_net_tables
    net = real_net_tables(agent)
    par = EnvParams(agent=agent, objective="queue", norm_wave=5.0, norm_wait=30.0, clip_wave=2.0, clip_wait=2.0,
                    coef_wait=0.0, coop_gamma=0.9, teleport_sec=300, real_net_norm=True, use_wait=False,
                    det_len=-1.0, halt_speed=0.1, queue_cap=10)
    n_step, reward_norm = 40, 1.0
else:
    net, par = build_large_grid(agent=agent), EnvParams(agent=agent)
    n_step, reward_norm = 120, 2000.0 if agent == "ma2c" else 3000.0
sim = BatchedSim(net, par, R)
lay = PolicyLayout(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, net.node_obs_off, net.n_obs, fw=128, ft=32,
                   ff=64 if agent == "ma2c" else 0, h=64, max_na=net.max_na, recurrent=policy != "fc")
if policy == "fc":
    from deeprl_signal_control_b200.agents.learner_fc import BatchedFcA2C as BatchedA2C
model = BatchedA2C(lay, R, n_step=n_step, gamma=0.99, v_coef=0.5, max_grad_norm=40.0, alpha=0.99, eps=1e-5,
                   reward_norm=reward_norm, reward_clip=2.0, seed=1, chunk=min(R, 1024))
tr = BatchedTrainer(sim, model, agent, lr=5e-4, beta=0.01, seed0=12)
t0 = time.time()
curve = []
while len(tr.episode_rewards) < episodes:
    tr.run(720)
    torch.cuda.synchronize()
    curve = list(tr.episode_rewards)
    print("episode %3d  mean step reward %9.2f   grad-norm[0] %.3f   %.1fs" %
          (len(curve), curve[-1], float(model.norms[0]), time.time() - t0), flush=True)
json.dump({"agent": agent, "scenario": scenario, "policy": policy, "replicas": R, "episodes": len(curve), "mean_episode_reward": curve,
           "wall_s": time.time() - t0, "env_steps": tr.n_env_steps}, open("gpurun_out/train_curve_%s_%s_%s.json" % (agent, scenario, policy), "w"))

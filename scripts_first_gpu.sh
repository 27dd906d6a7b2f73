set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch
from deeprl_signal_control_b200.net.large_grid import build_large_grid
from deeprl_signal_control_b200.net.tables import EnvParams
from deeprl_signal_control_b200.sim import BatchedSim
net, par = build_large_grid(agent='ma2c'), EnvParams(agent='ma2c')
for R in (1024, 8192):
    sim = BatchedSim(net, par, R)
    sim.reset(np.arange(R, dtype=np.uint64))
    print(sim.info())
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    fp = torch.rand(R, net.n_nodes, net.max_na, device='cuda')
    acts = [torch.randint(0, 5, (R, net.n_nodes), device='cuda', dtype=torch.int32) for _ in range(16)]
    for seg in range(6):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(60):
            sim.step(acts[i % 16], fp)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 60
        print(R, 'seg', seg, 'ms/step', ms, 'agent-env-steps/s', R * 25 / ms * 1e3, 'mean live', sim.mean_live())
    sim.close()
PY

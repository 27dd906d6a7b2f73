#!/usr/bin/env python
"""bench.py — agent-env-steps/sec of the 5x5 large_grid MA2C hot path on N B200s (one node).

Contract: `python bench.py --gpus N --steps K --warmup W` (N>1: launched by torchrun, one rank
per GPU).  Prints ONE JSON line on rank 0.

Workload (BASELINE.json configs[2]): 5x5 large_grid, MA2C observation layout (n_s in {32,42,52},
fingerprints), `--replicas` (default 8192) lock-stepped env replicas PER GPU ("weak" scaling),
synthetic demand of the named grid (large_grid/data/build_file.py flows 1100/925), replica r of
rank k seeded with seed0 + k*R + r.

A "step" is one control step (5 simulated seconds) of all local replicas:
   --mode sim   (default until the learner kernels land): uniform-random actions + fingerprints
                resident on the device -> tsc_step (one launch of tsc_step_kernel)
   --mode train: MA2C policy forward + sampling + tsc_step + transition store, and one n-step
                A2C update every n_step control steps (see deeprl_signal_control_b200.agents)
Before timing, every replica is advanced `--burnin` control steps (default 240 = 1200 simulated
seconds, the demand peak) so that the timed steps see a loaded network; burn-in is state
preparation, the W warm-up steps are on top of it.

value = (replicas over all ranks) * 25 agents * K / max-over-ranks(device time of K steps).
e2e   = same metric through the host-buffer entry point tsc_step_host (actions/fingerprints in
        pinned host memory -> H2D, kernel, obs/reward/done -> D2H, every step).
roofline = tsc_step_kernel: algorithmic bytes (BASELINE.md §3 formula with the measured mean
        live vehicles) / mean launch duration (CUDA events around every launch) vs measured HBM peak.
cpu_baseline / --impl reference = the CPU oracle (oracle/tsc_sim_ref.c, "port": SUMO is absent)
        on all host cores over a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = 25
N_STEP = 120          # batch_size of config/config_ma2c_large.ini


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=120)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--replicas", type=int, default=8192)
    p.add_argument("--burnin", type=int, default=240)
    p.add_argument("--mode", default=None, choices=[None, "sim", "train"])
    p.add_argument("--chunk", type=int, default=1024, help="replicas per BPTT chunk in the update")
    p.add_argument("--fp32-gemm", action="store_true", help="plain fp32 (no TF32 tensor cores) in the learner GEMMs")
    p.add_argument("--seed", type=int, default=12)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work of the oracle sample (both arms)")
    p.add_argument("--agent", default="ma2c", choices=["ma2c", "ia2c"],
                   help="ma2c = BASELINE configs[2] (the headline workload); ia2c with --policy fc = configs[1]")
    p.add_argument("--policy", default="lstm", choices=["lstm", "fc"], help="fc = FcACPolicy (agents/policies.py:214-256)")
    p.add_argument("--e2e-parts", type=int, default=3,
                   help="replica ranges of the host-buffer (e2e) loop, one stream each (1: single blocking tsc_step_host)")
    return p.parse_args()


def workload_name(R, mode, agent="ma2c", policy="lstm"):
    tag = "MA2C (configs[2])" if agent == "ma2c" else ("IA2C, FC policy (configs[1])" if policy == "fc" else "IA2C, LSTM policy")
    return ("5x5 large_grid %s, %d env replicas per GPU, %s" %
            (tag, R, "policy+sim+update" if mode == "train" else "sim control step, uniform-random actions"))


def algorithmic_bytes(net, v_live):
    """B_step = 2*V*16 + 2*L*8 + 2*A*4 + 4*A + 4*sum(N_s) + 4*A + 4 + 1."""
    L, A = net.n_lanes, net.n_nodes
    return 2 * v_live * 16 + 2 * L * 8 + 2 * A * 4 + 4 * A + 4 * net.n_obs + 4 * A + 4 + 1


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU every 100 ms while running."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                     0x4: "sw_power_cap", 0x80: "hw_power_brake"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.1)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference(net, par, args, threads, budget_s=12.0, quiet=True):
    """Time the CPU oracle on `threads` host threads over a bounded sample of the workload:
    R_cpu replicas, burn-in to the same simulated time, then timed control steps."""
    from oracle.sim_ref import RefSim
    rng = np.random.default_rng(0)
    # pilot: cost of one replica control step on one thread, at a lightly loaded network
    pilot = RefSim(net, par, threads)
    pilot.reset(np.arange(threads, dtype=np.uint64))
    for _ in range(40):
        pilot.step(rng.integers(0, 5, (threads, net.n_nodes), dtype=np.int32), None, threads=threads)
    t0 = time.perf_counter()
    for _ in range(40):
        pilot.step(rng.integers(0, 5, (threads, net.n_nodes), dtype=np.int32), None, threads=threads)
    c_step = (time.perf_counter() - t0) / 40            # seconds per (threads replicas) step
    n_t = int(min(max(args.steps, 60), 240))             # timed control steps (stay inside the episode)
    R_cpu = int(np.clip(budget_s / ((args.burnin + n_t) * c_step * 3.0) * threads, threads, 4096))
    R_cpu = max(threads, R_cpu // threads * threads)
    sim = RefSim(net, par, R_cpu)
    sim.reset(np.arange(R_cpu, dtype=np.uint64) + np.uint64(args.seed))
    fp = rng.random((R_cpu, net.n_nodes, net.max_na), dtype=np.float32)
    acts = rng.integers(0, 5, (8, R_cpu, net.n_nodes), dtype=np.int32)
    # one call per phase: every thread walks its replicas through all the steps (ref_run_mt), threads are created once
    sim.run(acts, args.burnin, fp, threads=threads)
    n, t0 = n_t, time.perf_counter()
    sim.run(acts, n_t, fp, threads=threads)
    el = time.perf_counter() - t0
    live = float(np.mean([sim.misc(r)["live"] for r in range(R_cpu)]))
    return {"value": R_cpu * net.n_nodes * n / el, "unit": "agent-env-steps/s", "cores": threads,
            "kind": "port",
            "sample": "%d replicas x %d control steps after %d burn-in steps (mean live %.0f veh/replica), "
                      "oracle/tsc_sim_ref.c on %d pthreads; SUMO+TF1 absent so this is the restatement, not SUMO"
                      % (R_cpu, n, args.burnin, live, threads)}, n, el, R_cpu


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    mode = args.mode or "train"
    from deeprl_signal_control_b200.net.large_grid import build_large_grid
    from deeprl_signal_control_b200.net.tables import EnvParams
    net, par = build_large_grid(agent=args.agent), EnvParams(agent=args.agent)
    cores = len(os.sched_getaffinity(0))

    # ---------------- reference arm: the CPU implementation of the path ----------------------
    if args.impl == "reference":
        if rank != 0:
            return
        t_steps = max(args.steps, 1)
        cb, n, el, R_cpu = cpu_reference(net, par, args, cores, budget_s=args.cpu_budget)
        line = {"impl": "reference", "metric": "agent-env-steps/sec", "value": cb["value"],
                "unit": "agent-env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * el / n, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.replicas, mode, args.agent, args.policy),
                           "note": "each reference step is a bounded sample: %d replicas instead of %d"
                                   % (R_cpu, args.replicas)},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "agent-env-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- our arm -----------------------------------------------------------------
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    from deeprl_signal_control_b200.sim import BatchedSim
    R = args.replicas
    sim = BatchedSim(net, par, R, device=local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches = 0
    trainer = None
    align_steps = 0
    if mode == "train":
        from deeprl_signal_control_b200.agents.layout import PolicyLayout
        from deeprl_signal_control_b200.agents.learner import BatchedA2C
        from deeprl_signal_control_b200.agents.trainer import BatchedTrainer
        # config/config_ma2c_large.ini [MODEL_CONFIG]
        lay = PolicyLayout(net.n_s_ls, net.n_a_ls, net.n_w_ls, net.n_f_ls, net.node_obs_off, net.n_obs,
                           fw=128, ft=32, ff=64 if args.agent == "ma2c" else 0, h=64, recurrent=args.policy != "fc")
        if args.policy == "fc":
            from deeprl_signal_control_b200.agents.learner_fc import BatchedFcA2C as Learner
        else:
            Learner = BatchedA2C
        model = Learner(lay, R, n_step=N_STEP, gamma=0.99, v_coef=0.5, max_grad_norm=40.0, alpha=0.99, eps=1e-5,
                        reward_norm=2000.0 if args.agent == "ma2c" else 3000.0, reward_clip=2.0, seed=args.seed,
                        device=local_rank, chunk=args.chunk, replica0=rank * R, total_replicas=world * R,
                        process_group=dist.group.WORLD if world > 1 else None, allow_tf32=not args.fp32_gemm)
        trainer = BatchedTrainer(sim, model, args.agent, lr=5e-4, beta=0.01, seed0=args.seed, replica0=rank * R)

        def one_step(i):
            trainer.control_step()
    else:
        seeds = np.arange(R, dtype=np.uint64) + np.uint64(args.seed + rank * R)
        sim.reset(seeds)
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)
        n_act_sets = 16
        acts = [torch.randint(0, 5, (R, net.n_nodes), device=dev, dtype=torch.int32, generator=gen)
                for _ in range(n_act_sets)]
        fp = torch.rand(R, net.n_nodes, net.max_na, device=dev, generator=gen)
        sim_events = []

        def one_step(i):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sim.step(acts[i % n_act_sets], fp)
            e1.record()
            sim_events.append((e0, e1))

    for i in range(args.burnin):
        one_step(i)
    for i in range(max(args.warmup, 3)):
        one_step(i)
    align_steps = 0
    if trainer is not None:
        # the timed region must never skip the learner update: align it so that it ENDS on an update boundary
        # (ceil(K / n_step) updates inside; for K < n_step this over-counts update work — conservative)
        while (trainer.model.t + args.steps) % N_STEP != 0:
            one_step(0)
            align_steps += 1
    live0 = sim.mean_live()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if trainer is not None:
        trainer.sim_events = []
        l0 = trainer.model.kernel_launches
        upd0 = trainer.n_updates
    else:
        sim_events.clear()
    barrier()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for i in range(args.steps):
        one_step(i)
    t_end.record()
    barrier()
    total_ms = t_start.elapsed_time(t_end)
    evs = trainer.sim_events if trainer is not None else sim_events
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    if trainer is not None:
        timed_launches = args.steps + (trainer.model.kernel_launches - l0)
        n_updates_timed = trainer.n_updates - upd0
        trainer.sim_events = None
    else:
        timed_launches = args.steps
        n_updates_timed = 0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    live1 = sim.mean_live()
    v_live = 0.5 * (live0 + live1)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * R * net.n_nodes * args.steps / (total_ms_max * 1e-3)

    # ---------------- e2e: the environment driven through the host-buffer C-ABI call ------------
    if trainer is not None:
        e2e_steps = N_STEP                     # one full rollout + one update
        if args.policy == "fc":
            args.e2e_parts = 1                 # the replica-range forward exists for the fused LSTM kernel only
        if args.e2e_parts > 1:
            host_step = lambda: trainer.control_step_host_pipelined(n_parts=args.e2e_parts)
        else:
            host_step = trainer.control_step_host
        for i in range(3):
            host_step()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            host_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        h2d = R * net.n_nodes * 4 + R * net.n_nodes * net.max_na * 4 + R * net.n_obs * 4 + R * net.n_nodes * 4 + R * 4
        d2h = R * net.n_nodes * 4 + R * net.n_nodes * net.max_na * 4 + R * net.n_obs * 4 + R * net.n_nodes * 4 + R * 4 + R
        e2e_api = ("BatchedTrainer.control_step_host: policy forward on device, actions+fingerprints D2H, "
                   "tsc_step_host (H2D, kernel, D2H), obs+reward H2D, update every 120 steps")
        if args.e2e_parts > 1:
            e2e_api = ("BatchedTrainer.control_step_host_pipelined: %d replica ranges, one stream each; per range: policy "
                       "forward (tscl_policy_step_v2r), actions+fingerprints D2H to pinned host buffers, "
                       "tsc_step_host_range (H2D, kernel, D2H, host sync), obs+reward H2D into the learner; update "
                       "every 120 steps" % args.e2e_parts)
    else:
        e2e_steps = max(3, min(args.steps, 20))
        h_act = [torch.randint(0, 5, (R, net.n_nodes), dtype=torch.int32).pin_memory().numpy() for _ in range(4)]
        h_fp = torch.rand(R, net.n_nodes, net.max_na).pin_memory().numpy()
        sim._h_out = tuple(torch.from_numpy(a).pin_memory().numpy() for a in (
            np.zeros((R, net.n_obs), np.float32), np.zeros((R, net.n_nodes), np.float32),
            np.zeros(R, np.float32), np.zeros(R, np.uint8)))
        for i in range(3):
            sim.step_host(h_act[i % 4], h_fp)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            sim.step_host(h_act[i % 4], h_fp)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        h2d = R * net.n_nodes * 4 + R * net.n_nodes * net.max_na * 4
        d2h = R * net.n_obs * 4 + R * net.n_nodes * 4 + R * 4 + R
        e2e_api = "tsc_step_host (pinned host buffers)"
    t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * R * net.n_nodes * e2e_steps / (float(t.item()) * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel ----------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = algorithmic_bytes(net, v_live) * R
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tr_path = os.path.join(ROOT, "profiles", "r01_sim_kernel_traffic.json")
    if os.path.exists(tr_path):
        traffic = json.load(open(tr_path)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "tsc_step_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "mean_live_vehicles_per_replica": v_live,
                "kernel_ms_per_launch": kern_ms, "share_of_step": kern_ms * args.steps / total_ms,
                "note": "kernel is issue-bound, not HBM-bound: ~1.2k instructions per vehicle-second x5 "
                        "fused sub-steps per 32 B of state traffic (DESIGN.md §5)"}
    cb = None
    if not args.no_cpu_baseline:
        cb, _, _, _ = cpu_reference(net, par, args, cores, budget_s=args.cpu_budget)
    line = {"metric": "agent-env-steps/sec", "value": value, "unit": "agent-env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if mode == "sim" else "f32 (sim, LSTM cell, loss, optimizer) + bf16 tensor-core operands with f32 accumulation (learner GEMMs)",
            "data": "synthetic",
            "config": {"workload": workload_name(R, mode, args.agent, args.policy), "replicas_per_gpu": R, "agents": net.n_nodes,
                       "burnin_control_steps": args.burnin, "mode": mode, "n_step": N_STEP,
                       "updates_in_timed_region": n_updates_timed, "untimed_alignment_steps": align_steps,
                       "learner_gemm_library": "own tcgen05 kernels for the forward, BPTT and all weight gradients; cuBLAS "
                                               "bf16 only for dX = dZ.Wx^T (1 plain batched GEMM per update chunk)"
                       if mode == "train" else None,
                       "l2": "inputs larger than L2: %.0f MB of replica state per GPU is streamed every step"
                             % (R * sim.info()["state_bytes_per_replica"] / 1e6),
                       "parallelism": "replica-dp%d" % world},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "agent-env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": e2e_api},
            "gpu_launches": timed_launches,
            "roofline": roofline, "cpu_baseline": cb}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

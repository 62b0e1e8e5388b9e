#!/usr/bin/env python
"""bench.py -- SamCon hot-path benchmark (contract: see the task's Measurement section / DESIGN.md §7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one SamCon window of BASELINE.json configs[1] ("walk.yml full nIter SamCon solve, nSample scaled to
10k per window", synthetic SMPL-shaped reference motion): draw nSample=10 000 perturbed PD targets, roll every
sample 100 x {stable PD, stepSimulation} from its parent elite, score, keep the nSave=1000 best, which become the
next step's parents (windows are chained exactly like the reference's solve).  With N GPUs every rank carries
10 000 samples / 1 000 elites of a global 10 000*N / 1 000*N search (weak scaling) and elites are exchanged per
window (samcon_b200/distributed.py).

value  = sample-sim-steps/s, whole job, everything resident in HBM (actions drawn on the device).
e2e    = same metric through the pool's host API (what SamCon.sample calls): per step the (S,68) actions come
         from pinned host memory and (state, cost, info) of every sample go back to the host, like the
         reference's pipe scatter/gather.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

S_PER_GPU, E_PER_GPU, STEPS_PER_WINDOW = 10000, 1000, 100
FLOP_PER_SAMPLE_STEP = 2.5e5      # F(n_c=4)
BYTES_PER_SAMPLE_STEP = 21.0
METRIC, UNIT = "humanoid sample-sim-steps/sec", "sample-steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum of one rollout launch of this workload (S = 10 000, 100 steps), from the
# committed `ncu --set full` capture; below the algorithmic 21 MB because one window's buffers sit in the 126 MB L2
NCU_DRAM_BYTES_PER_LAUNCH = 6747648 + 881152
NCU_DRAM_SOURCE = "profiles/r1s_walk_rollout_full.txt (ncu --set full, one launch, S=10000)"


def workload_inputs(n_windows):
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    cfg = Config("walk")
    ch = Character.from_config(cfg)
    motion = make_synthetic_motion("gait", T=max(100, 3 * n_windows + 10), fr=30, seed=0)
    (name, _), = motion.items()
    loader = AmassMotionData(motion, name)
    targets = loader.all_window_states(n_windows, 1.0 / cfg.fps_act)
    return cfg, ch, targets


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None
        return self

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.p:
            self.p.terminate()
            try:
                self.p.wait(timeout=2)
            except Exception:
                self.p.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_throughput(ch, targets, n_samples, cores, steps=STEPS_PER_WINDOW, reps=1):
    """The CPU restatement on `cores` host threads (ctypes releases the GIL): one window of n_samples rollouts from
    the bench's own start state, split contiguously like the reference splits elites over workers
    (samcon.py:85-90).  -> (sample-steps/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.sampling import get_batch_action
    cfg = Config("walk")
    o = Oracle(ch)
    E = max(1, n_samples // 10)
    parents = np.repeat(targets[0][None], E, axis=0)
    actions = get_batch_action(targets[1], E * 10, cfg.values, "uniform", np.random.RandomState(cfg.seed))
    S = E * 10
    bounds = np.linspace(0, S, cores + 1).astype(int)
    t0 = time.perf_counter()
    for _ in range(reps):
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(lambda i: o.window(parents, actions, targets[1], steps, int(bounds[i]), int(bounds[i + 1])), range(cores)))
    dt = time.perf_counter() - t0
    return S * steps * reps / dt, dt, S


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as omod
    omod.build()
    cores = os.cpu_count() or 1
    cfg, ch, targets = workload_inputs(2)
    n = cores * 24
    for _ in range(args.warmup):
        cpu_oracle_throughput(ch, targets, max(10, cores * 2), cores)
    t0 = time.perf_counter()
    tot = 0
    for _ in range(args.steps):
        v, dt, S = cpu_oracle_throughput(ch, targets, n, cores)
        tot += S * STEPS_PER_WINDOW
    el = time.perf_counter() - t0
    val = tot / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "walk.yml SamCon window, nSample 10k / nSave 1k per GPU, 100 sim steps per window, "
                                   "synthetic SMPL-shaped walk; CPU arm: bounded sample per step", "sample_per_step": S},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{S} rollouts x {STEPS_PER_WINDOW} steps per step on {cores} host threads; CPU "
                                       "restatement (oracle/samcon_oracle.c), not PyBullet -- pybullet 3.2.5 is not "
                                       "installable here"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from samcon_b200 import build as b
    if rank == 0:
        b.build()
    if world > 1:
        dist.barrier()
    from samcon_b200.distributed import ShardedSearch
    from samcon_b200.pool import GpuSamplePool
    from samcon_b200.sampling import get_batch_action

    K, W = args.steps, args.warmup
    S, E = args.samples, args.samples // 10
    cfg, ch, targets = workload_inputs(W + K + 1)
    pool = GpuSamplePool(ch, local)
    search = ShardedSearch(pool, S, E * world, rank, world)
    d_targets = torch.tensor(targets, dtype=torch.float32, device=dev)
    limits = torch.tensor(cfg.values, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident arm
    parents = d_targets[0:1].repeat(E, 1).contiguous()
    kern_events = []

    def device_step(w, parents, timed):
        flush.zero_()                                           # L2 flush between steps (inside the timed bracket)
        act = pool.sample_gen(d_targets[w], limits, S, seed=cfg.seed + 1000 * rank, window=w)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        blk, ids, ecost, outs = None, None, None, None
        st, cost, info = pool.window(parents, act, d_targets[w], STEPS_PER_WINDOW)
        if timed:
            e1.record()
            kern_events.append((e0, e1))
        elites, ids, ecost = search.select_and_exchange(st, cost)
        return elites[rank * E:(rank + 1) * E].contiguous(), ecost

    for w in range(1, W + 1):
        parents, ecost = device_step(w, parents, False)
    sync_all()
    l0 = pool.launch_count
    with ClockSampler(local) as clk:
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for w in range(W + 1, W + K + 1):
            parents, ecost = device_step(w, parents, True)
        t1.record()
        sync_all()
    launches = pool.launch_count - l0
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_events) / K], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = ms.item(), kms.item()
    best_cost = float(ecost[0].item())
    clocks = clk.summary()

    # ------------------------------------------------------------------ end-to-end arm (host API, pinned host inputs)
    rs = np.random.RandomState(cfg.seed + rank)
    host_actions = [torch.from_numpy(get_batch_action(targets[w], S, cfg.values, cfg.noise_type, rs).astype(np.float32)).pin_memory()
                    for w in range(1, W + K + 1)]
    pool.init(targets[0], E)
    sink = 0.0

    def e2e_step(w):
        st, cost, info = pool.sim(targets[w], host_actions[w - 1], STEPS_PER_WINDOW)      # H2D actions, D2H results
        inds = pool.select_last(E)
        pool.keep(inds)
        return float(cost[inds[0]])
    for w in range(1, W + 1):
        sink += e2e_step(w)
    sync_all()
    c0 = time.perf_counter()
    for w in range(W + 1, W + K + 1):
        sink += e2e_step(w)
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - c0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_val = world * S * STEPS_PER_WINDOW * K / e2e_s.item()

    # ------------------------------------------------------------------ BASELINE metric (ii): seconds per SamCon window at
    # walk.yml's own size (nSample 2000, nSave 200), through the host API like the e2e arm; rank 0, N = 1 only
    walk_window_s = None
    if world == 1:
        S2, E2 = cfg.nSample, cfg.nSave
        acts2 = [torch.from_numpy(get_batch_action(targets[w], S2, cfg.values, cfg.noise_type, rs).astype(np.float32)).pin_memory()
                 for w in range(1, W + K + 1)]
        pool.init(targets[0], E2)

        def walk_step(w):
            st, cost, info = pool.sim(targets[w], acts2[w - 1], STEPS_PER_WINDOW)
            pool.keep(pool.select_last(E2))
        for w in range(1, W + 1):
            walk_step(w)
        torch.cuda.synchronize(dev)
        c0 = time.perf_counter()
        for w in range(W + 1, W + K + 1):
            walk_step(w)
        torch.cuda.synchronize(dev)
        walk_window_s = (time.perf_counter() - c0) / K

    value = world * S * STEPS_PER_WINDOW * K / (ms_total * 1e-3)
    kern_rate = S * STEPS_PER_WINDOW / (kernel_ms * 1e-3)       # one GPU's rollout kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    roofline = {"bound": "fp32", "bound_note": "FP32 CUDA-core pipe; neither hbm (0.004 % of DRAM peak) nor tensor (no dense "
                                                  "contraction in a 57-DoF rigid-body step): SURVEY.md §8d, DESIGN.md §4",
                "achieved": kern_rate * FLOP_PER_SAMPLE_STEP / 1e12,
                "peak": fp32_peak, "unit": "TFLOP/s", "frac": kern_rate * FLOP_PER_SAMPLE_STEP / 1e12 / fp32_peak,
                "traffic": NCU_DRAM_BYTES_PER_LAUNCH, "traffic_source": NCU_DRAM_SOURCE,
                "algorithmic_bytes_per_launch": S * STEPS_PER_WINDOW * BYTES_PER_SAMPLE_STEP,
                "kernel": "sc_rollout_kernel", "kernel_ms_per_launch": kernel_ms,
                "flop_per_sample_step": FLOP_PER_SAMPLE_STEP,
                "peak_source": f"nominal 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json"
                               f"{'' if peaks else ' absent: fallback 1965'})",
                "hbm": {"achieved": kern_rate * BYTES_PER_SAMPLE_STEP / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": kern_rate * BYTES_PER_SAMPLE_STEP / 1e9 / hbm_peak,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "walk.yml SamCon solve windows, nSample 10k / nSave 1k per GPU, 100 sim steps per window "
                                   "(dt 1 ms, 2 sub-steps, 10 solver iters), synthetic SMPL-shaped walk",
                       "nSample_per_gpu": S, "nSave_per_gpu": E, "steps_per_window": STEPS_PER_WINDOW,
                       "l2": "256 MB memset between steps (inside the timed region)", "parallelism": f"samples sharded x{world}",
                       "best_elite_cost_last_window": best_cost,
                       "walk_yml_window_s_at_nSample_2000": walk_window_s},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": S * 68 * 4 + 132 * 4,
                    "d2h_bytes_per_step": S * (132 + 1 + 5) * 8 + E * 4},     # float64 results (widened on the device), int32 elite ids
            "roofline": roofline}
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v, dt, Sc = cpu_oracle_throughput(ch, targets, cores * 800, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{Sc} rollouts x {STEPS_PER_WINDOW} steps of the same window in {dt:.1f} s on {cores} "
                                          "host threads; CPU restatement (oracle), not PyBullet"}
    if rank == 0:
        print(json.dumps(line))
    pool.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=S_PER_GPU, help="nSample per GPU (default: configs[1]'s 10k)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

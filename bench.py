#!/usr/bin/env python
"""bench.py -- SamCon hot-path benchmark (contract: see the task's Measurement section / DESIGN.md §7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config walk|sweep|cartwheel|multiseq]

A *step* is one SamCon window: draw the perturbed PD targets, roll every sample 100 x {stable PD, stepSimulation} from
its parent elite (state + contact cache), score, keep the best tenth, which become the next step's parents (windows
are chained exactly like the reference's solve, /root/reference/samcon/core/samcon.py:158-183).

--config (BASELINE.json `configs`):
  walk      #2, the default and the driver's line: walk.yml solve, nSample 10 000 / nSave 1 000 PER GPU (weak scaling), samples
            sharded over the GPUs with one elite exchange per window (samcon_b200/distributed.py)
  sweep     #4: 10^6 rollouts per window IN TOTAL, split over the GPUs (strong scaling)
  cartwheel #3: cartwheel.yml (hand/foot ground contact), 10 000 / 1 000 per GPU, sharded like `walk`
  multiseq  #5: 64 synthetic sequences x walk.yml's own 2 000 / 200 solved concurrently, sequences dealt to the GPUs (no collective)
All on synthetic SMPL-shaped reference motion (the planted-foot gait / the fold-to-the-hands motion; no AMASS data ships).

value  = sample-sim-steps/s, whole job, everything resident in HBM (actions drawn on the device).
e2e    = same metric through the host-facing API: per step the (S,68) actions come from pinned host memory and (state, cost,
         info) of every sample go back to the host as float64, like the reference's pipe scatter/gather -- at N > 1 through
         the same sharded search, collective included.
"""
from __future__ import annotations

import argparse
import glob
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

STEPS_PER_WINDOW = 100
FLOP_PER_SAMPLE_STEP = 2.5e5      # F(n_c=4)
BYTES_PER_SAMPLE_STEP = 21.0
METRIC, UNIT = "humanoid sample-sim-steps/sec", "sample-steps/s"

CONFIGS = {
    # name: (cfg, motion kind, samples per GPU (None: total / N), total samples (None: per GPU x N), scaling, BASELINE config)
    "walk": ("walk", "gait", 10000, None, "weak", 2),
    "sweep": ("walk", "gait", None, 1000000, "strong", 4),
    "cartwheel": ("cartwheel", "touchdown", 10000, None, "weak", 3),
    "multiseq": ("walk", "gait", None, None, "strong", 5),
}
WORKLOAD = {
    "walk": "walk.yml SamCon solve windows (BASELINE config #2), nSample 10k / nSave 1k per GPU, 100 sim steps per window (dt 1 ms, "
            "2 sub-steps, 10 solver iters, warm-started contacts), chained windows of the synthetic SMPL-shaped planted-foot gait",
    "sweep": "throughput sweep (BASELINE config #4): 10^6 rollouts x 0.1 s window in total, split over the GPUs, walk.yml settings, "
             "chained windows of the synthetic planted-foot gait",
    "cartwheel": "cartwheel.yml SamCon solve windows (BASELINE config #3: contact-rich, hand and foot ground contact), nSample 10k / "
                 "nSave 1k per GPU sharded with the per-window elite exchange, chained windows of the synthetic fold-to-the-hands "
                 "motion (the kinematic cartwheel cannot be followed by any search)",
    "multiseq": "multi-sequence batching (BASELINE config #5): 64 synthetic gait sequences x nSample 2000 / nSave 200 solved "
                "concurrently, sequences dealt to the GPUs, one launch per window per GPU",
}
MULTISEQ_Q, MULTISEQ_S, MULTISEQ_E = 64, 2000, 200


def static_config(name, world, samples=0):
    """The `config` object of the JSON line: what the workload is, nothing that depends on the run -- both arms print the
    same one (the reference arm times a bounded sample of it, described in its cpu_baseline.sample)."""
    cfgname, kind, per_gpu, total, scaling, bcfg = CONFIGS[name]
    if name == "multiseq":
        return {"workload": WORKLOAD[name], "baseline_config": bcfg, "sequences": MULTISEQ_Q, "nSample": MULTISEQ_S, "nSave": MULTISEQ_E,
                "steps_per_window": STEPS_PER_WINDOW, "l2": "256 MB memset between steps (inside the timed region)",
                "parallelism": f"sequences dealt x{world}"}
    S = samples or (per_gpu if per_gpu else total // world)
    S -= S % 10
    return {"workload": WORKLOAD[name], "baseline_config": bcfg, "nSample_per_gpu": S, "nSave_per_gpu": S // 10,
            "steps_per_window": STEPS_PER_WINDOW, "l2": "256 MB memset between steps (inside the timed region)",
            "parallelism": f"samples sharded x{world}"}


def workload_inputs(n_windows, cfgname="walk", kind="gait", seed=0):
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    cfg = Config(cfgname)
    ch = Character.from_config(cfg)
    # long enough that no timed window's target wraps around the end of the cycle
    motion = make_synthetic_motion(kind, T=max(100, 3 * n_windows + 10), fr=30, seed=seed)
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


# ----------------------------------------------------------------------------------------------------------------------
# CPU side: the oracle port on all host threads (ctypes releases the GIL)
# ----------------------------------------------------------------------------------------------------------------------
def oracle_window(o, parents, pcache, actions, target, nsteps, cores):
    """One window of len(actions) rollouts split contiguously over `cores` threads like the reference splits elites over
    its workers (samcon.py:85-90).  -> (state, cost, info, cache)"""
    from concurrent.futures import ThreadPoolExecutor
    S = len(actions)
    bounds = np.linspace(0, S, cores + 1).astype(int)
    st = np.zeros((S, 132)); cost = np.zeros(S); info = np.zeros((S, 5)); cache = np.zeros((S, o.cache_dim))

    def job(i):
        a, b = int(bounds[i]), int(bounds[i + 1])
        if b > a:
            st[a:b], cost[a:b], info[a:b], cache[a:b] = o.window(parents, actions, target, nsteps, a, b, parent_cache=pcache,
                                                                  return_cache=True)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(job, range(cores)))
    return st, cost, info, cache


def run_reference(args):
    """The reference arm: the CPU implementation of the path on the box's host cores.  pybullet==3.2.5 (the reference's
    physics) is not installable here, so this is the oracle port (cpu_baseline.kind = "port"), driven like the reference's
    pool: a SamCon search of its own on the same motion with chained windows (elites of window t are the parents of t+1,
    contact caches included), each step one window of a bounded number of rollouts."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as omod
    omod.build()
    from oracle.oracle import Oracle
    from samcon_b200.sampling import get_batch_action
    cfgname, kind, _, _, scaling, _ = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    cfg, ch, targets = workload_inputs(W + K + 1, cfgname, kind)
    o = Oracle(ch)
    S = cores * 48
    S -= S % 10
    E = S // 10
    rng = np.random.RandomState(cfg.seed)
    parents, pcache = np.repeat(targets[0][None], E, axis=0), None
    t0 = None
    for w in range(1, W + K + 1):
        if w == W + 1:
            t0 = time.perf_counter()
        acts = get_batch_action(targets[w], S, cfg.values, cfg.noise_type, rng)
        st, cost, info, cache = oracle_window(o, parents, pcache, acts, targets[w], STEPS_PER_WINDOW, cores)
        idx = np.lexsort((np.arange(S), cost))[:E]
        parents, pcache = st[idx], cache[idx]
    el = time.perf_counter() - t0
    val = S * STEPS_PER_WINDOW * K / el
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": 1e3 * el / K, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": static_config(args.config, args.gpus, args.samples),
            "solve": {"sample_per_step": S, "best_elite_cost_last_window": float(cost[idx[0]])},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"a search of its own over the same motion, chained windows, {S} rollouts x {STEPS_PER_WINDOW} "
                                       f"steps per step on {cores} host threads; CPU restatement (oracle/samcon_oracle.c, float64), "
                                       "not PyBullet -- pybullet 3.2.5 is not installable here"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------------------------------
def rollout_traffic():
    """DRAM bytes of one rollout launch from this round's committed ncu capture (tools/ncu_summary.py output), or None."""
    # profiles/LATEST_ROLLOUT_CAPTURE names the capture of the committed kernel (written together with it); else the last by name
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2*_rollout_full.txt")))
    latest = os.path.join(ROOT, "profiles", "LATEST_ROLLOUT_CAPTURE")
    if os.path.exists(latest):
        named = os.path.join(ROOT, "profiles", open(latest).read().strip())
        if os.path.exists(named):
            files.append(named)
    if not files:
        return None, None
    rd = wr = None
    for line in open(files[-1]):
        p = line.split()
        if len(p) >= 2 and p[0] == "dram__bytes_read.sum":
            rd = float(p[1].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(p[2] if len(p) > 2 else "byte", 1)
        if len(p) >= 2 and p[0] == "dram__bytes_write.sum":
            wr = float(p[1].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(p[2] if len(p) > 2 else "byte", 1)
    if rd is None or wr is None:
        return None, None
    return rd + wr, os.path.relpath(files[-1], ROOT) + " (ncu --set full, one sc_rollout_kernel launch of this workload)"


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
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
    if args.config == "multiseq":
        return run_multiseq(args, world, rank, local, dev)
    from samcon_b200 import abi
    from samcon_b200.distributed import ShardedSearch
    from samcon_b200.pool import GpuSamplePool
    from samcon_b200.sampling import get_batch_action

    cfgname, kind, per_gpu, total, scaling, bcfg = CONFIGS[args.config]
    K, W = args.steps, args.warmup
    S = args.samples or (per_gpu if per_gpu else total // world)
    S -= S % 10
    E = S // 10
    cfg, ch, targets = workload_inputs(W + K + 2, cfgname, kind)
    pool = GpuSamplePool(ch, local)
    search = ShardedSearch(pool, S, E * world, rank, world)
    d_targets = torch.tensor(targets, dtype=torch.float32, device=dev)
    limits = torch.tensor(cfg.values, dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cost = torch.empty((S,), dtype=torch.float32, device=dev)
    info = torch.empty((S, 5), dtype=torch.float32, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident arm
    kern_events = []

    def device_step(w, parents, pcache, timed):
        flush.zero_()                                           # L2 flush between steps (inside the timed bracket)
        act = pool.sample_gen(d_targets[w], limits, S, seed=cfg.seed + 1000 * rank, window=w)
        st, cache = search.result_buffers()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        pool.window(parents, act, d_targets[w], STEPS_PER_WINDOW, out=(st, cost, info), parent_cache=pcache, out_cache=cache)
        if timed:
            e1.record()
            kern_events.append((e0, e1))
        bst, bcache, ids, ecost = search.select_and_exchange(st, cost, cache)
        return bst, bcache, ids, ecost, act, st, cache

    parents, pcache = d_targets[0:1].repeat(E, 1).contiguous(), None
    check = None
    for w in range(1, W + 1):
        prev = (parents, pcache)
        parents, pcache, ids, ecost, act, st, cache = device_step(w, parents, pcache, False)
        if w == W and world > 1:
            # on-hardware proof of the exchange (outside the timed region): the global top-E recomputed from ALL ranks'
            # costs equals the exchanged ids, and this rank's parent block holds the owners' rows bit for bit
            all_cost = torch.empty((world * S,), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(all_cost, cost.contiguous())
            all_st = torch.empty((world * S, 132), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(all_st, st.contiguous())
            all_ch = torch.empty((world * S, cache.shape[1]), dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(all_ch, cache.contiguous())
            # (all_gathered rows are in (rank, local) order; ids are in the unsharded search's numbering)
            lr = torch.arange(world * S, device=dev)
            order = search.global_sample_ids(lr // S, lr % S)
            ref_pos = torch.argsort(all_cost, stable=True)[:E * world]
            # ties between ranks aside (measure zero), the E cheapest samples, cheapest first
            ok = torch.equal(order[ref_pos], ids) and torch.equal(all_st[ref_pos[rank::world]], parents) \
                and torch.equal(all_ch[ref_pos[rank::world]], pcache)
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            check = "ok" if flag.item() == 1 else "MISMATCH"
    pool.contact_stats(reset=True)
    sync_all()
    l0 = pool.launch_count
    with ClockSampler(local) as clk:
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for w in range(W + 1, W + K + 1):
            prev = (parents, pcache)
            parents, pcache, ids, ecost, act, st, cache = device_step(w, parents, pcache, True)
        t1.record()
        sync_all()
    launches = pool.launch_count - l0
    overflow, substeps, overflow_touching = pool.contact_stats()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_events) / K], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = ms.item(), kms.item()
    best_cost = float(ecost[0].item())
    clocks = clk.summary()
    # inputs of the last timed window, for the CPU leg below (same parents, caches, actions, target)
    cpu_in = (prev[0].cpu().numpy().astype(np.float64), prev[1].cpu().numpy().astype(np.float64) if prev[1] is not None else None,
              act.cpu().numpy().astype(np.float64), cost.cpu().numpy().astype(np.float64)) if (rank == 0 and world == 1 and not args.no_cpu) else None

    # ------------------------------------------------------------------ end-to-end arm (host buffers in, host results out)
    rs = np.random.RandomState(cfg.seed + rank)
    host_actions = [torch.from_numpy(get_batch_action(targets[w], S, cfg.values, cfg.noise_type, rs).astype(np.float32)).pin_memory()
                    for w in range(1, W + K + 1)]
    stage = (torch.empty((S, 132), dtype=torch.float64).pin_memory(), torch.empty((S,), dtype=torch.float64).pin_memory(),
             torch.empty((S, 5), dtype=torch.float64).pin_memory(), torch.empty((E * world,), dtype=torch.int64).pin_memory())
    h_target = [torch.from_numpy(targets[w].astype(np.float32)).pin_memory() for w in range(W + K + 1)]
    state = {"parents": d_targets[0:1].repeat(E, 1).contiguous(), "pcache": None}
    sink = 0.0

    def e2e_step(w):
        act = host_actions[w - 1].to(dev, non_blocking=True)                    # H2D: this step's actions and target
        tgt = h_target[w].to(dev, non_blocking=True)
        st, cache = search.result_buffers()
        pool.window(state["parents"], act, tgt, STEPS_PER_WINDOW, out=(st, cost, info), parent_cache=state["pcache"], out_cache=cache)
        bst, bcache, ids, ecost = search.select_and_exchange(st, cost, cache)
        state["parents"], state["pcache"] = bst, bcache
        for hbuf, d in zip(stage, (st, cost, info)):                            # D2H: every sample's result, float64 like the
            hbuf.copy_(d.double(), non_blocking=True)                            # reference's result tuples, + the elite ids
        stage[3].copy_(ids, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return float(stage[1][int(stage[3][0]) % S]) if world == 1 else float(stage[1][0])
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
    # walk.yml's own size (nSample 2000, nSave 200), through the host protocol (pool.sim / select_last / keep); N = 1 only
    walk_window_s = None
    if world == 1 and args.config == "walk":
        S2, E2 = cfg.nSample, cfg.nSave
        acts2 = [get_batch_action(targets[w], S2, cfg.values, cfg.noise_type, rs).astype(np.float32) for w in range(1, W + K + 1)]
        pool.init(targets[0], E2)

        def walk_step(w):
            pool.sim(targets[w], acts2[w - 1], STEPS_PER_WINDOW)
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
    fp32_nominal = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    import ctypes as C
    tf = C.c_double(0.0)
    fp32_measured = tf.value if abi.load().sc_measure_fp32_peak(local, C.byref(tf)) == 0 else 0.0
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic, traffic_src = rollout_traffic()
    achieved = kern_rate * FLOP_PER_SAMPLE_STEP / 1e12
    roofline = {"bound": "fp32", "bound_note": "FP32 CUDA-core pipe; neither hbm (~0.01 % of DRAM peak) nor tensor (no dense "
                                                  "contraction in a 57-DoF rigid-body step): SURVEY.md §8d, DESIGN.md §4",
                "achieved": achieved, "peak": fp32_nominal, "unit": "TFLOP/s", "frac": achieved / fp32_nominal,
                "peak_measured_fp32": fp32_measured or None,
                "frac_of_measured_peak": achieved / fp32_measured if fp32_measured else None,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": S * STEPS_PER_WINDOW * BYTES_PER_SAMPLE_STEP,
                "kernel": "sc_rollout_kernel", "kernel_ms_per_launch": kernel_ms,
                "flop_per_sample_step": FLOP_PER_SAMPLE_STEP,
                "peak_source": f"nominal 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json"
                               f"{'' if peaks else ' absent: fallback 1965'}), which has no FP32 figure; peak_measured_fp32 = "
                               "sc_measure_fp32_peak (register-only FFMA loop on all SMs, this run)",
                "hbm": {"achieved": kern_rate * BYTES_PER_SAMPLE_STEP / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": kern_rate * BYTES_PER_SAMPLE_STEP / 1e9 / hbm_peak,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": static_config(args.config, world, args.samples),
            # what this run found (not part of the workload's description)
            "solve": {"elite_exchange": search.exchange, "exchange_check": check,
                      "best_elite_cost_last_window": best_cost, "solve_tracks_motion": bool(best_cost < 10.0),
                      "contact_cap": {"max_contacts": 8, "frac_substeps_over_cap_in_margin": overflow / max(1, substeps),
                                      "frac_substeps_over_cap_touching": overflow_touching / max(1, substeps)},
                      "walk_yml_window_s_at_nSample_2000": walk_window_s},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": S * 68 * 4 + 132 * 4,
                    "d2h_bytes_per_step": S * (132 + 1 + 5) * 8 + E * world * 8},   # float64 results (widened on the device), int64 elite ids
            "roofline": roofline}
    if cpu_in is not None:
        # the CPU port on the very window the GPU just simulated: same parents, contact caches, actions and target; a bounded
        # number of its samples (whole families of an elite's children)
        from oracle.oracle import Oracle
        cores = os.cpu_count() or 1
        o = Oracle(ch)
        n_cpu = min(S, cores * 800)
        n_cpu -= n_cpu % 10
        P, PC, A, gcost = cpu_in
        if PC is None:
            PC = o.empty_cache(len(P))
        c0 = time.perf_counter()
        ost, ocost, _, _ = oracle_window(o, P[: n_cpu // 10], PC[: n_cpu // 10], A[:n_cpu], targets[W + K], STEPS_PER_WINDOW, cores)
        dt = time.perf_counter() - c0
        rel = np.abs(gcost[:n_cpu] - ocost) / np.abs(ocost)
        line["cpu_baseline"] = {"value": n_cpu * STEPS_PER_WINDOW / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"the first {n_cpu} rollouts x {STEPS_PER_WINDOW} steps of the last timed window (the GPU's own "
                                          f"parents, contact caches and actions) in {dt:.1f} s on {cores} host threads; CPU restatement "
                                          "(oracle), not PyBullet",
                                "same_window_cost_agreement": {"median_rel": float(np.median(rel)), "frac_within_1e-3": float((rel < 1e-3).mean())}}
    if rank == 0:
        print(json.dumps(line))
    pool.close()
    if world > 1:
        dist.destroy_process_group()


def run_multiseq(args, world, rank, local, dev):
    """BASELINE config 5: 64 sequences solved concurrently; sequences, not samples, are dealt to the GPUs (no collective)."""
    import torch
    import torch.distributed as dist
    from samcon_b200.batched import BatchedSamCon
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    from samcon_b200.pool import GpuSamplePool
    K, W = args.steps, args.warmup
    cfg = Config("walk")
    pool = GpuSamplePool(Character.from_config(cfg), local)
    Q = MULTISEQ_Q
    per = (Q + world - 1) // world
    q0, q1 = min(Q, rank * per), min(Q, (rank + 1) * per)
    loaders = []
    for q in range(q0, q1):
        m = make_synthetic_motion("gait", T=max(100, 3 * (W + K) + 10), seed=q)
        (name, _), = m.items()
        loaders.append(AmassMotionData(m, name))
    b = BatchedSamCon(cfg, loaders, pool, nIter=W + K, nSample=MULTISEQ_S, nSave=MULTISEQ_E, seed_offset=q0)
    b.begin()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
    for t in range(1, W + 1):
        b.step(t)
    sync_all()
    l0 = pool.launch_count
    with ClockSampler(local) as clk:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for t in range(W + 1, W + K + 1):
            flush.zero_()
            ecost = b.step(t)
        t1.record()
        sync_all()
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    worst = torch.tensor([float(ecost[:, 0].max().item())], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    value = Q * MULTISEQ_S * STEPS_PER_WINDOW * K / (ms.item() * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms.item() / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": static_config("multiseq", world),
            "solve": {"sequences_per_gpu": q1 - q0, "sequence_windows_per_s": Q * K / (ms.item() * 1e-3),
                      "worst_best_elite_cost_last_window": worst.item()},
            "clocks": clk.summary(), "gpu_launches": int(pool.launch_count - l0)}
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
    ap.add_argument("--config", default="walk", choices=sorted(CONFIGS))
    ap.add_argument("--samples", type=int, default=0, help="override nSample per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

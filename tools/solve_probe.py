#!/usr/bin/env python
"""Developer probe: run a SamCon solve on the CPU oracle (all host threads) and print per-window solve quality.

    python tools/solve_probe.py [--cfg walk] [--motion walk] [--nsample 2000] [--nsave 200] [--niter 30] [--seed 1]

Prints, per window: best / median / worst elite cost, the best elite's root height against the target's, and the
cost terms of the best elite.  Used to find out why the round-1 bench workload fell over.
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_window(o, parents, actions, target, nsteps, cores):
    S = len(actions)
    bounds = np.linspace(0, S, cores + 1).astype(int)
    st = np.zeros((S, 132)); cost = np.zeros(S); info = np.zeros((S, 5))

    def job(i):
        a, b = int(bounds[i]), int(bounds[i + 1])
        if b > a:
            st[a:b], cost[a:b], info[a:b] = o.window(parents, actions, target, nsteps, a, b)
    with ThreadPoolExecutor(cores) as ex:
        list(ex.map(job, range(cores)))
    return st, cost, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="walk")
    ap.add_argument("--motion", default="walk")
    ap.add_argument("--nsample", type=int, default=2000)
    ap.add_argument("--nsave", type=int, default=200)
    ap.add_argument("--niter", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--sim", default="", help="SimParams overrides, e.g. erp=0.2,num_solver_iters=10")
    ap.add_argument("--save", default="", help="npz to store the best path in")
    args = ap.parse_args()
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.model import Character, SimParams
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion as make_motion
    from samcon_b200.sampling import get_batch_action
    cfg = Config(args.cfg)
    kw = {}
    for item in filter(None, args.sim.split(",")):
        k, v = item.split("=")
        kw[k] = type(getattr(SimParams(), k))(float(v)) if not isinstance(getattr(SimParams(), k), bool) else bool(int(v))
    sim = SimParams(dt=1.0 / cfg.fps_sim, self_collision=bool(cfg.self_collision), **kw)
    ch = Character.from_config(cfg, sim=sim)
    o = Oracle(ch)
    motion = make_motion(args.motion)
    (name, _), = motion.items()
    loader = AmassMotionData(motion, name)
    dt = 1.0 / cfg.fps_act
    nsteps = cfg.fps_sim // cfg.fps_act
    rng = np.random.RandomState(args.seed)
    S, E = args.nsample, args.nsave
    cores = os.cpu_count() or 1
    parents = np.repeat(loader.getSpecTimeState(0.0)[None], E, axis=0)
    hist = []
    t00 = time.time()
    for t in range(1, args.niter + 1):
        target = loader.getSpecTimeState(t * dt)
        acts = get_batch_action(target, S, cfg.values, cfg.noise_type, rng)
        st, cost, info = oracle_window(o, parents, acts, target, nsteps, cores)
        idx = np.lexsort((np.arange(S), cost))[:E]
        parents = st[idx]
        b = idx[0]
        hist.append((st[idx], acts[idx], cost[idx], idx // (S // E)))
        print(f"w{t:02d} elite cost {cost[idx[0]]:8.3f} {np.median(cost[idx]):8.3f} {cost[idx[-1]]:8.3f} | worst {cost.max():8.2f} | "
              f"root z {st[b, 2]:.3f} (target {target[2]:.3f}) y {st[b, 1]:.3f} ({target[1]:.3f}) | terms " +
              " ".join(f"{cfg.cost_weights[i] * info[b, i]:.3f}" for i in range(5)) + f" | {time.time() - t00:.0f}s", flush=True)
    if args.save:
        e = 0
        sim_, act_ = [], []
        for P, A, Cc, Ep in reversed(hist):
            sim_.append(P[e]); act_.append(A[e]); e = int(Ep[e])
        np.savez(args.save, state=np.array(sim_[::-1]), action=np.array(act_[::-1]), t0=loader.getSpecTimeState(0.0))


if __name__ == "__main__":
    main()

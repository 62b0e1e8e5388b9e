"""The five BASELINE.json configs on one GPU, device-resident (BatchedSamCon: no host round trip inside a solve):
    python tools/config_report.py > profiles/<round>_configs.md
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from samcon_b200.batched import BatchedSamCon
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.motion import AmassMotionData, make_synthetic_motion
from samcon_b200.pool import GpuSamplePool


def loaders(kind, n):
    out = []
    for seed in range(n):
        m = make_synthetic_motion(kind, T=91, seed=seed)
        (name, _), = m.items()
        out.append(AmassMotionData(m, name))
    return out


def solve(cfgname, kind, Q, S, E, nIter):
    cfg = Config(cfgname)
    pool = GpuSamplePool(Character.from_config(cfg), 0)
    b = BatchedSamCon(cfg, loaders(kind, Q), pool, nIter=nIter, nSample=S, nSave=E)
    b.nIter = 2; b.sample(); b.nIter = nIter                      # warm-up (allocator, kernels)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = b.sample()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    pool.close()
    return dt, float(np.mean([r["elite_cost"][:, 0].max() for r in res]))


rows = []
dt, c = solve("walk", "gait", 1, 2000, 200, 30)
rows.append(("#1 walk.yml, nSample 2000 / nSave 200", "30 windows", f"{dt:.3f} s per solve, {dt/30*1e3:.1f} ms per window, {2000*100*30/dt:.3e} sample-steps/s; worst best-elite cost {c:.2f}"))
dt, c = solve("walk", "gait", 1, 10000, 1000, 30)
rows.append(("#2 walk.yml full solve, nSample 10 000 / nSave 1000", "30 windows", f"{dt:.3f} s per solve, {dt/30*1e3:.1f} ms per window, {10000*100*30/dt:.3e} sample-steps/s; worst best-elite cost {c:.2f}"))
dt, c = solve("cartwheel", "touchdown", 1, 2000, 200, 30)
rows.append(("#3 cartwheel.yml on the fold-to-the-hands motion (one GPU's share), nSample 2000 / nSave 200", "30 windows", f"{dt:.3f} s per solve, {dt/30*1e3:.1f} ms per window; worst best-elite cost {c:.2f}"))
dt, c = solve("walk", "gait", 1, 1000000, 100000, 3)
rows.append(("#4 10^6 rollouts x 0.1 s window (one GPU)", "3 windows", f"{dt/3:.3f} s per window, {1e6*100*3/dt:.3e} sample-steps/s"))
dt, c = solve("walk", "gait", 64, 2000, 200, 30)
rows.append(("#5 64 sequences x nSample 2000 concurrently (one launch per window)", "30 windows", f"{dt:.3f} s for all 64 solves, {64*30/dt:.1f} sequence-windows/s, {64*2000*100*30/dt:.3e} sample-steps/s; mean over sequences of the worst best-elite cost {c:.2f}"))
print("| Config | Work | One B200, device resident |\n|---|---|---|")
for r in rows:
    print("| " + " | ".join(r) + " |")

"""Developer probe (GPU): how robustly does a walk.yml-size solve track variants of the planted-foot gait?
    python tools/gait_sweep.py [nSample nSave]
For every variant Q = 4 independent searches (different Philox streams) run concurrently through BatchedSamCon; prints the
worst best-elite cost over the 30 windows and the worst root-height error of the best path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from samcon_b200.batched import BatchedSamCon
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.motion import AmassMotionData, make_gait_motion
from samcon_b200.pool import GpuSamplePool

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
E = int(sys.argv[2]) if len(sys.argv) > 2 else S // 10
Q = 4
cfg = Config("walk")
pool = GpuSamplePool(Character.from_config(cfg), 0)
base = dict(step_time=0.5, step_length=0.4, double_support=0.2, pelvis_height=0.93, half_width=0.09, sway=0.04, lift=0.06, ramp=0.3)
variants = [("base", {})]
for k, vals in (("step_length", (0.25, 0.3, 0.5)), ("sway", (0.0, 0.02, 0.06)), ("pelvis_height", (0.91, 0.95)), ("step_time", (0.4, 0.6, 0.7)),
                ("double_support", (0.1, 0.3, 0.4)), ("half_width", (0.07, 0.11, 0.13)), ("lift", (0.04, 0.09))):
    for v in vals:
        variants.append((f"{k}={v}", {k: v}))
variants += [("slow", dict(step_time=0.6, step_length=0.3, double_support=0.3)), ("slow2", dict(step_time=0.7, step_length=0.3, double_support=0.3, sway=0.06))]
print(f"nSample {S} nSave {E}, {Q} searches per variant")
for name, kw in variants:
    m = make_gait_motion(T=100, **{**base, **kw})
    (mn, _), = m.items()
    ld = AmassMotionData(m, mn)
    b = BatchedSamCon(cfg, [ld] * Q, pool, nIter=30, nSample=S, nSave=E)
    res = b.sample()
    worst = [float(r["elite_cost"][:, 0].max()) for r in res]
    dz = [float(np.abs(r["simulated_state"][:, 2] - r["target_state"][:, 2]).max()) for r in res]
    tot = [r["total_cost"] for r in res]
    print(f"{name:22s} worst best-elite cost " + " ".join(f"{w:6.2f}" for w in worst) + " | dz " + " ".join(f"{d:.3f}" for d in dz) +
          " | path total " + " ".join(f"{t:6.1f}" for t in tot), flush=True)
pool.close()

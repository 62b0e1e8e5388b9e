"""Developer probe (GPU): can a search follow the synthetic cartwheel at all?  Per-window best elite cost / root height for a
few sample counts; optionally stores the best path of the largest one as a physically feasible motion (50 fps).
    python tools/cartwheel_probe.py [out.npz]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from samcon_b200.batched import BatchedSamCon
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.motion import AmassMotionData, make_synthetic_motion
from samcon_b200.pool import GpuSamplePool

cfg = Config("cartwheel")
pool = GpuSamplePool(Character.from_config(cfg), 0)
m = make_synthetic_motion("cartwheel", T=100, seed=0)
(name, _), = m.items()
ld = AmassMotionData(m, name)
for S in (2000, 20000, 200000):
    b = BatchedSamCon(cfg, [ld], pool, nIter=30, nSample=S, nSave=S // 10)
    r = b.sample()[0]
    z = r["simulated_state"][:, 2]; zt = r["target_state"][:, 2]
    print(f"S={S}: path total {r['total_cost']:.1f}; per-window best elite cost " + " ".join(f"{c:.1f}" for c in r["elite_cost"][:, 0]))
    print("   root z sim/target: " + " ".join(f"{a:.2f}/{b_:.2f}" for a, b_ in zip(z[::3], zt[::3])))
pool.close()

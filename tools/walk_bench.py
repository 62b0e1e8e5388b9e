"""Developer timing probe: bench.py's chained walk.yml windows (device sample-gen, select, gather), per-window kernel time.
    [SC_LIB=libsamcon_b200_x.so] [SC_SELF=0] python tools/walk_bench.py [S] [windows]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from samcon_b200 import abi
if os.environ.get("SC_LIB"):
    abi.LIB_PATH = os.path.join(ROOT, "samcon_b200", "csrc", os.environ["SC_LIB"])
import bench
from samcon_b200.config import Config
from samcon_b200.model import Character, SimParams
from samcon_b200.pool import GpuSamplePool
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = Config("walk")
ch = Character.from_config(cfg, SimParams(self_collision=os.environ.get("SC_SELF", "1") == "1",
                                           exact_box_contacts=os.environ.get("SC_EXACT_BOX", "0") == "1"))
pool = GpuSamplePool(ch, 0)
_, _, targets = bench.workload_inputs(W + 1)
d_targets = torch.tensor(targets, dtype=torch.float32, device="cuda")
limits = torch.tensor(cfg.values, dtype=torch.float32, device="cuda")
E = S // 10
parents, pcache = d_targets[0:1].repeat(E, 1).contiguous(), None
ms = []
for w in range(1, W + 1):
    act = pool.sample_gen(d_targets[w], limits, S, seed=cfg.seed, window=w)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st, cost, info, cache = pool.window(parents, act, d_targets[w], 100, parent_cache=pcache, out_cache=True); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
    idx, ec = pool.select(cost, E)
    parents, pcache = pool.gather_rows(st, idx.reshape(-1)), pool.gather_rows(cache, idx.reshape(-1))
print(os.environ.get("SC_LIB", "default"), "self" if ch.sim.self_collision else "noself", f"S={S}",
      " ".join(f"{m:.1f}" for m in ms), f"| mean of last {W - 2}: {np.mean(ms[2:]):.2f} ms  best elite cost {ec[0, 0].item():.3f}")

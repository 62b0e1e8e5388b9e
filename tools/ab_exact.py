"""Developer A/B: two builds of the library on the same chained walk.yml windows -- outputs must be bit-identical (the
change under test only moves who computes what), per-window kernel times side by side.
    python tools/ab_exact.py libsamcon_b200_noshare.so[,libsamcon_b200_x.so...] [S] [windows]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from samcon_b200 import abi
import bench
from samcon_b200.config import Config
from samcon_b200.model import Character, SimParams
from samcon_b200.pool import GpuSamplePool

others = sys.argv[1].split(",")
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
W = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = Config("walk")
ch = Character.from_config(cfg, SimParams(self_collision=True))
default_lib = abi.LIB_PATH
pools = {}
for name, path in [("default", default_lib)] + [(o, os.path.join(ROOT, "samcon_b200", "csrc", o)) for o in others]:
    abi.LIB_PATH = path
    abi._lib = None          # load() caches the library: drop it so the next pool binds the other build
    pools[name] = GpuSamplePool(ch, 0)
_, _, targets = bench.workload_inputs(W + 1)
d_targets = torch.tensor(targets, dtype=torch.float32, device="cuda")
limits = torch.tensor(cfg.values, dtype=torch.float32, device="cuda")
E = S // 10
parents, pcache = d_targets[0:1].repeat(E, 1).contiguous(), None
ms = {k: [] for k in pools}
ref = pools["default"]
for w in range(1, W + 1):
    act = ref.sample_gen(d_targets[w], limits, S, seed=cfg.seed, window=w)
    outs = {}
    for k, pool in pools.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = pool.window(parents, act, d_targets[w], 100, parent_cache=pcache, out_cache=True); e1.record()
        torch.cuda.synchronize()
        ms[k].append(e0.elapsed_time(e1)); outs[k] = o
    a = outs["default"]
    same = [all(bool(torch.equal(x.view(torch.int32), y.view(torch.int32))) for x, y in zip(a, outs[o])) for o in others]
    print(f"window {w}: state/cost/info/cache bit-identical to default: {same}  ms " + "  ".join(f"{k} {v[-1]:.2f}" for k, v in ms.items()))
    assert all(same)
    st, cost, info, cache = a
    idx, ec = ref.select(cost, E)
    parents, pcache = ref.gather_rows(st, idx.reshape(-1)), ref.gather_rows(cache, idx.reshape(-1))
for k in ms:
    print(f"{k}: mean of last {W - 2} windows {np.mean(ms[k][2:]):.3f} ms")

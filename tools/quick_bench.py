"""Developer timing probe (not the contract bench): rollout kernel throughput for a few sample counts."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.pool import GpuSamplePool
from helpers import stand, near_action, rand_state

from samcon_b200.model import SimParams
cfg = Config("walk"); ch = Character.from_config(cfg, SimParams(self_collision=os.environ.get("SC_SELF", "1") == "1")); pool = GpuSamplePool(ch, 0)
print(pool.rollout_config())
rng = np.random.default_rng(0)
for mode in ("stand", "air"):
    for S in [int(x) for x in (sys.argv[1:] or [2000, 10000, 100000])]:
        E = max(1, S // 10)
        base = np.array([stand(rng, 0.03, 0.1, 0.985) if mode == "stand" else rand_state(rng, 0.3, 3.0) for _ in range(64)])
        parents = torch.tensor(base[rng.integers(0, 64, E)], dtype=torch.float32, device="cuda")
        target = torch.tensor(stand(rng), dtype=torch.float32, device="cuda")
        acts = torch.tensor(np.array([near_action(base[0], rng, 0.1) for _ in range(256)])[rng.integers(0, 256, S)], dtype=torch.float32, device="cuda")
        out = pool.window(parents, acts, target, 100)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            pool.window(parents, acts, target, 100, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{mode} S={S}: {ms:.2f} ms/window  {S*100/ms*1e3:.3e} sample-steps/s  cost mean {out[1].mean().item():.3f}")

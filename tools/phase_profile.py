"""Developer probe: per-phase cycle shares of the rollout kernel (build with -DSC_PROFILE into a separate .so).
    python tools/phase_profile.py [S]
"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from samcon_b200 import abi, build as b
lib = os.path.join(ROOT, "samcon_b200", "csrc", "libsamcon_b200_prof.so")
flags = [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
subprocess.check_call(["/usr/local/cuda/bin/nvcc", *flags, "-DSC_PROFILE", "-o", lib, os.path.join(ROOT, "samcon_b200", "csrc", "samcon_b200.cu")])
abi.LIB_PATH = lib
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.pool import GpuSamplePool
from helpers import stand, near_action, rand_state
NAMES = ["load", "fk", "pd_term", "aba_pd", "collide", "aba", "fk_vel", "c_rhs", "c_assembly", "c_gs", "c_apply", "integrate",
         "features+cost+store", "-", "-", "idle/other"]
cfg = Config("walk"); ch = Character.from_config(cfg); pool = GpuSamplePool(ch, 0)
pool._lib.sc_prof_read.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
rng = np.random.default_rng(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
for mode in ("stand", "air"):
    E = S // 10
    base = np.array([stand(rng, 0.03, 0.1, 0.985) if mode == "stand" else rand_state(rng, 0.3, 3.0) for _ in range(64)])
    parents = torch.tensor(base[rng.integers(0, 64, E)], dtype=torch.float32, device="cuda")
    target = torch.tensor(stand(rng), dtype=torch.float32, device="cuda")
    acts = torch.tensor(np.array([near_action(base[0], rng, 0.1) for _ in range(256)])[rng.integers(0, 256, S)], dtype=torch.float32, device="cuda")
    out = pool.window(parents, acts, target, 100)
    buf = (C.c_longlong * 16)()
    pool._lib.sc_prof_read(pool._h, buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pool.window(parents, acts, target, 100, out=out); e1.record(); torch.cuda.synchronize()
    pool._lib.sc_prof_read(pool._h, buf)
    v = np.array(list(buf), dtype=np.float64); tot = v.sum()
    print(f"== {mode} S={S}: {e0.elapsed_time(e1):.2f} ms")
    for n, x in zip(NAMES, v):
        if x: print(f"  {n:22s} {100*x/tot:6.2f}%")

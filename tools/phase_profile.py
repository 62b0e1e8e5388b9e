"""Developer probe: per-phase cycle shares of the rollout kernel (build with -DSC_PROFILE into a separate .so).
    python tools/phase_profile.py [S] [walk|stand|air ...]
`walk` replays bench.py's chained walk.yml windows (device sample-gen, cfg noise, select, gather).
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
         "features+cost+store", "cta_barrier_wait", None, "idle/other"]
from samcon_b200.model import SimParams
cfg = Config("walk")
ch = Character.from_config(cfg, SimParams(self_collision=os.environ.get("SC_SELF", "1") == "1", num_solver_iters=int(os.environ.get("SC_NITER", "10"))))
pool = GpuSamplePool(ch, 0)
pool._lib.sc_prof_read.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
rng = np.random.default_rng(0)
args = sys.argv[1:]
S = int(args[0]) if args and args[0].isdigit() else 10000
modes = [a for a in args if not a.isdigit()] or ["stand", "air"]
buf = (C.c_longlong * (40 + 2048))()


def report(tag, ms):
    pool._lib.sc_prof_read(pool._h, buf)
    v = np.array(list(buf), dtype=np.float64)
    ph, st = v[:16], v[16:32]
    tot = ph.sum() - ph[14] + v[26:31].sum()
    print(f"== {tag}: {ms:.2f} ms")
    for n, x in zip(NAMES, ph):
        if x and n: print(f"  {n:22s} {100*x/tot:6.2f}%  {ms*x/tot:7.3f} ms")
    for n, i in (("wait before collide (after aba)", 30), ("wait after collide", 26), ("wait after assembly", 27), ("wait after gs", 28), ("wait after apply", 29)):
        if v[i]: print(f"  {n:32s} {100*v[i]/tot:6.2f}%")
    if v[39]:
        # shared assembly (cycles after the barrier in front of collide): mean over the warps / mean of the CTA's maximum
        print(f"  assembly pool, cycles after the collide barrier, mean over warps | mean of the CTA maximum: tile ready {v[32]/st[0]:.0f} | {v[35]/v[39]:.0f}, "
              f"own calls done {v[33]/st[0]:.0f} | {v[36]/v[39]:.0f}, left the pool {v[34]/st[0]:.0f} | {v[37]/v[39]:.0f}; "
              f"per warp-substep: {(int(v[14]) & 0xffffff) / st[0]:.2f} own calls in {v[25]/st[0]:.0f} cycles, {(int(v[14]) >> 24) / st[0]:.2f} calls of other tiles in {v[31]/st[0]:.0f} cycles")
    ntr = int(v[40])
    if ntr and os.environ.get("SC_TRACE"):
        # the calls of CTA 0 in a few sub-steps: epoch, warp, tile, call, two-body tile, start and end (cycles after the collide barrier)
        ev = sorted((int(v[48 + 4 * k]), int(v[49 + 4 * k]), int(v[50 + 4 * k]), int(v[51 + 4 * k])) for k in range(min(ntr, 400)))
        for ep, code, t0_, t1_ in ev:
            print(f"    epoch {ep} warp {code & 255:2d} tile {code >> 8 & 255} call {code >> 16 & 255} {'two-body' if code >> 24 else '        '} {t0_:6d} .. {t1_:6d}  ({t1_ - t0_})")
    if st[0]:
        # per warp-substep statistics: st[0] substeps, st[1] with contacts, st[2..6] ncmax buckets 0,1-2,3-4,5-6,7-8,
        # st[7] overflow (some sample had more candidate hits than max_contacts), st[8] sum of per-sample nc (lane 0's)
        print(f"  warp-substeps {st[0]:.0f}: ncmax 0:{st[2]/st[0]:.2%} 1-2:{st[3]/st[0]:.2%} 3-4:{st[4]/st[0]:.2%} "
              f"5-6:{st[5]/st[0]:.2%} 7-8:{st[6]/st[0]:.2%}; overflow {st[7]/st[0]:.2%}; mean nc (sample 0 of tile) {st[8]/st[0]:.2f}")


for mode in modes:
    E = S // 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if mode == "walk":
        sys.path.insert(0, ROOT)
        import bench
        _, _, targets = bench.workload_inputs(9)
        d_targets = torch.tensor(targets, dtype=torch.float32, device="cuda")
        limits = torch.tensor(cfg.values, dtype=torch.float32, device="cuda")
        parents, pcache = d_targets[0:1].repeat(E, 1).contiguous(), None
        for w in range(1, 9):
            act = pool.sample_gen(d_targets[w], limits, S, seed=cfg.seed, window=w)
            pool._lib.sc_prof_read(pool._h, buf)
            e0.record(); st, cost, info, cache = pool.window(parents, act, d_targets[w], 100, parent_cache=pcache, out_cache=True); e1.record(); torch.cuda.synchronize()
            report(f"walk window {w} S={S}", e0.elapsed_time(e1))
            idx, ec = pool.select(cost, E)
            parents, pcache = pool.gather_rows(st, idx.reshape(-1)), pool.gather_rows(cache, idx.reshape(-1))
            print(f"  elite cost best {ec[0,0].item():.3f} worst {ec[0,-1].item():.3f}; all-sample mean {cost.mean().item():.3f}")
        continue
    base = np.array([stand(rng, 0.03, 0.1, 0.985) if mode == "stand" else rand_state(rng, 0.3, 3.0) for _ in range(64)])
    parents = torch.tensor(base[rng.integers(0, 64, E)], dtype=torch.float32, device="cuda")
    target = torch.tensor(stand(rng), dtype=torch.float32, device="cuda")
    acts = torch.tensor(np.array([near_action(base[0], rng, 0.1) for _ in range(256)])[rng.integers(0, 256, S)], dtype=torch.float32, device="cuda")
    out = pool.window(parents, acts, target, 100)
    pool._lib.sc_prof_read(pool._h, buf)
    e0.record(); pool.window(parents, acts, target, 100, out=out); e1.record(); torch.cuda.synchronize()
    report(f"{mode} S={S}", e0.elapsed_time(e1))

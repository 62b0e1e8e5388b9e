"""Developer probe (GPU): sc_select timings and equality with a stable argsort at the sizes of BASELINE.json's configs.
    python tools/select_probe.py > profiles/<round>_select_probe.log"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from samcon_b200.config import Config
from samcon_b200.model import Character
from samcon_b200.pool import GpuSamplePool
pool = GpuSamplePool(Character.from_config(Config("walk")), 0)
for S, E, Q in ((2000, 200, 1), (10000, 1000, 1), (8000, 1000, 1), (64000, 8000, 1), (1000000, 100000, 1), (128000, 200, 64), (640000, 1000, 64), (1000000, 1000, 1)):
    c = torch.rand(S, device="cuda")
    seg = (torch.arange(Q + 1, device="cuda", dtype=torch.int32) * (S // Q)).contiguous()
    pool.select(c, E, seg_offsets=seg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        idx, ec = pool.select(c, E, seg_offsets=seg)
    e1.record(); torch.cuda.synchronize()
    ref = torch.argsort(c.reshape(Q, -1), dim=1, stable=True)[:, :E] + seg[:-1, None].long()
    print(f"select S={S} E={E} segments={Q}: {e0.elapsed_time(e1) / 10:.3f} ms, equal to argsort: {torch.equal(idx.long(), ref)}")
# non-finite inputs do not hang sample generation
t = torch.zeros(132, device="cuda"); t[10::4][:17] = 1
lim = torch.full((51,), float("inf"), device="cuda")
a = pool.sample_gen(t, lim, 64, seed=1, window=1); torch.cuda.synchronize(); print("sample_gen with inf limits returned", a.shape)

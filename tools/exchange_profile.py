"""Developer probe (torchrun, N GPUs): where a sharded window's time goes besides the rollout kernel.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_profile.py [windows]
Per rank: CUDA-event times of rollout / local select / all_gather / final select / peer fetch, and the wait a rank spends
in the all_gather for slower ranks (the per-window straggler effect of a synchronous elite selection)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import bench
from samcon_b200.distributed import ShardedSearch
from samcon_b200.pool import GpuSamplePool

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 12
S, E = 10000, 1000
cfg, ch, targets = bench.workload_inputs(W + 2)
pool = GpuSamplePool(ch, local)
search = ShardedSearch(pool, S, E * world, rank, world)
d_targets = torch.tensor(targets, dtype=torch.float32, device=dev)
limits = torch.tensor(cfg.values, dtype=torch.float32, device=dev)
cost = torch.empty((S,), dtype=torch.float32, device=dev); info = torch.empty((S, 5), dtype=torch.float32, device=dev)
parents, pcache = d_targets[0:1].repeat(E, 1).contiguous(), None
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for w in range(1, W + 1):
    act = pool.sample_gen(d_targets[w], limits, S, seed=cfg.seed + 1000 * rank, window=w)
    st, cache = search.result_buffers()
    e = [ev() for _ in range(6)]
    e[0].record()
    pool.window(parents, act, d_targets[w], 100, out=(st, cost, info), parent_cache=pcache, out_cache=cache)
    e[1].record()
    e[2].record()
    all_cost = torch.empty((world * S,), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_cost, cost.contiguous())
    e[3].record()
    pos, gcost = pool.select(all_cost, search.E)
    pos = pos.reshape(-1)
    win_rank = torch.div(pos, S, rounding_mode="floor").to(torch.int32)
    win_local = (pos - win_rank * S).to(torch.int32).contiguous()
    e[4].record()
    parents, pcache = search.fetch_block(win_rank, win_local, st, cache)
    search._win += 1
    e[5].record()
    torch.cuda.synchronize()
    rows.append([e[i].elapsed_time(e[i + 1]) for i in range(5)])
r = np.array(rows[3:])
m = torch.tensor(r.mean(0), device=dev)
allm = [torch.zeros_like(m) for _ in range(world)]
dist.all_gather(allm, m)
if rank == 0:
    print(f"{world} GPUs, {len(r)} windows, ms per window per rank: rollout | - | all_gather of the costs (incl. waiting "
          f"for the slowest rank) | select ({search.E} of {world * S}) + index math | peer fetch ({E} x 592 B)")
    for k, a in enumerate(allm):
        print(f"rank {k}: " + " | ".join(f"{x:7.3f}" for x in a.tolist()))
    kern = torch.stack(allm)[:, 0]
    print(f"rollout mean over ranks {kern.mean().item():.3f}, max {kern.max().item():.3f}")
pool.close(); dist.destroy_process_group()

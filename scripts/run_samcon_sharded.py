"""One SamCon solve with the samples of every window sharded over the GPUs of one box (one process per GPU):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_samcon_sharded.py --cfg cartwheel \\
        [--nSample 16000 --nSave 1600] [--no-save]

The multi-GPU form of scripts/run_samcon.py (reference: one process pool per run, /root/reference/scripts/run_samcon.py:19-40):
samcon_b200.distributed.ShardedSamCon -- rollouts need no communication, one elite exchange per window (all_gather of
(cost, index), winners fetched over NVLink).  Rank 0 prints the summary and, unless --no-save, writes the best path as
<output_dir>/<cfg>_<timestamp>_sharded.pkl (path_0 of the reference's result layout)."""
import argparse
import datetime
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch                                                  # noqa: E402
import torch.distributed as dist                              # noqa: E402

from samcon_b200.config import Config                       # noqa: E402
from samcon_b200.distributed import ShardedSamCon           # noqa: E402
from samcon_b200.model import Character                     # noqa: E402
from samcon_b200.motion import AmassMotionData, make_synthetic_motion  # noqa: E402
from samcon_b200.pool import GpuSamplePool                  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=str, default="walk")
    ap.add_argument("--nSample", type=int, default=None, help="samples per window of the whole search (default: cfg.nSample x ranks)")
    ap.add_argument("--nSave", type=int, default=None)
    ap.add_argument("--no-save", action="store_true")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = Config(args.cfg, base_dir="", create_dirs=not args.no_save and rank == 0)
    if os.path.exists(cfg.data_path):
        data = AmassMotionData(cfg.data_path, cfg.motion_seq)
        if cfg.auto_nIter:
            cfg.nIter = int(data.getCycleTime() * cfg.fps_act)
    else:
        motion = make_synthetic_motion("touchdown" if "cartwheel" in args.cfg else "gait", T=3 * cfg.nIter + 10, fr=30, seed=0)
        (name, _), = motion.items()
        cfg.motion_seq = name
        data = AmassMotionData(motion, name)
    S = args.nSample or cfg.nSample * world
    E = args.nSave or S // (cfg.nSample // cfg.nSave)
    pool = GpuSamplePool(Character.from_config(cfg), local)
    algo = ShardedSamCon(cfg, data, pool, rank, world, nSample=S, nSave=E)
    torch.cuda.synchronize()
    t0 = time.time()
    res = algo.sample()
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"cfg {args.cfg}: {cfg.nIter} windows x {S} samples on {world} GPU(s) ({algo.search.exchange}) in {dt:.2f} s "
              f"({dt / cfg.nIter:.3f} s per window), best path total cost {res['total_cost']:.3f}, worst best-elite cost "
              f"{res['elite_cost'][:, 0].max():.2f}")
        if not args.no_save:
            import joblib
            stamp = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S")
            out = {cfg.motion_seq: {"path_0": {k: res[k] for k in ("target_state", "simulated_state", "action", "cost", "total_cost", "t0_state")}}}
            joblib.dump(out, os.path.join(cfg.output_dir, f"{cfg.id}_{stamp}_sharded.pkl"))
    pool.close()
    if world > 1:
        dist.destroy_process_group()

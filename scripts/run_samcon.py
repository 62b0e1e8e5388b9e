#!/usr/bin/env python
"""Run one SamCon solve on the GPU path -- the counterpart of the reference's ``scripts/run_samcon.py``
(/root/reference/scripts/run_samcon.py:14-40): same ``--cfg`` / ``--num_processes`` arguments, same flow
(Config -> AmassMotionData -> auto_nIter -> SamCon(...).sample()), results written in the reference's pickle layout
under ``results/samcon/<cfg>/``.  ``--num_processes`` is accepted and ignored (one GPU pool replaces the worker
pool).  When ``cfg.data_path`` does not exist (the AMASS-derived pickle is not shipped), ``--synthetic`` solves the
bundled synthetic SMPL-shaped motion of the same kind instead.

    python scripts/run_samcon.py --cfg walk [--synthetic] [--nSample 10000 --nSave 1000] [--no-save]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from samcon_b200.config import Config                       # noqa: E402
from samcon_b200.motion import AmassMotionData, make_synthetic_motion  # noqa: E402
from samcon_b200.samcon import SamCon                       # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=str, default="walk")
    ap.add_argument("--num_processes", type=int, default=12)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--nSample", type=int, default=None)
    ap.add_argument("--nSave", type=int, default=None)
    ap.add_argument("--no-save", action="store_true")
    args = ap.parse_args()

    cfg = Config(args.cfg, base_dir="", create_dirs=not args.no_save)
    if args.synthetic or not os.path.exists(cfg.data_path):
        if not args.synthetic:
            print(f"{cfg.data_path} not found: solving the synthetic '{args.cfg}' motion (pass --synthetic to silence)")
        # generated long enough that the last window's target (t = nIter / fps_act) lies inside the cycle
        motion = make_synthetic_motion("touchdown" if "cartwheel" in args.cfg else "gait", T=3 * cfg.nIter + 10, fr=30, seed=0)
        (name, _), = motion.items()
        cfg.motion_seq = name
        data = AmassMotionData(motion, name)
        cfg.auto_nIter = False
    else:
        data = AmassMotionData(cfg.data_path, cfg.motion_seq)
    if cfg.auto_nIter:
        cfg.nIter = int(data.getCycleTime() * cfg.fps_act)
    algo = SamCon(cfg, data, args.num_processes, cfg.nIter, args.nSample or cfg.nSample, args.nSave or cfg.nSave)
    t0 = time.time()
    out = algo.sample(save=not args.no_save)
    dt = time.time() - t0
    algo.close()
    res = out[0] if isinstance(out, tuple) else out
    best = res[cfg.motion_seq]["path_0"]["total_cost"] if isinstance(res, dict) and cfg.motion_seq in res else float("nan")
    print(f"cfg {args.cfg}: {cfg.nIter} windows x {args.nSample or cfg.nSample} samples in {dt:.2f} s "
          f"({dt / max(1, cfg.nIter):.3f} s per window), best path total cost {best:.3f}")

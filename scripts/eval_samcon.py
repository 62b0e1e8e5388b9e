"""Headless counterpart of the reference's scripts/eval_samcon.py (/root/reference/scripts/eval_samcon.py:22-81):
load a result pickle written by run_samcon.py, replay the best path through the env facade -- ``_sim_agent.set_state(t0)``,
then per window ``_kin_agent.set_state(target); env.step(action)`` -- and report how closely the replay reproduces the
stored trajectory (on the GPU path: bit for bit) and the per-window costs.  The GUI animation and the matplotlib cost
scatter plots of the reference script are outside this repo's scope.

    python scripts/eval_samcon.py --cfg walk --file walk_<timestamp>.pkl
"""
import argparse
import os
import os.path as osp
import sys

import numpy as np

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))

from samcon_b200.config import Config                       # noqa: E402
from samcon_b200.envs import HumanoidTrackingEnv            # noqa: E402


def replay(env, path):
    """-> (simulated states (n,132), costs (n,)) of replaying one result path through env.step."""
    env._sim_agent.set_state(path["t0_state"])
    states, costs = [], []
    for i in range(path["action"].shape[0]):
        env._kin_agent.set_state(path["target_state"][i])
        s, c, _ = env.step(path["action"][i])
        states.append(s); costs.append(c)
    return np.array(states), np.array(costs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", type=str, default="walk")
    ap.add_argument("--file", type=str, required=True, help="file name under cfg.output_dir (as written by run_samcon.py)")
    args = ap.parse_args()
    import joblib
    cfg = Config(args.cfg, base_dir="", create_dirs=False)
    m_data = joblib.load(osp.join(cfg.output_dir, args.file))
    i_path = osp.join(cfg.info_dir, args.file)
    if osp.exists(i_path):
        i_data = joblib.load(i_path)
        print(f"info: {i_data['total_cost'].shape[0]} windows x {i_data['total_cost'].shape[1]} samples, "
              f"{i_data['elite_inds'].shape[1]} elites; best path's elite slots {i_data['best_path_my_sys'].tolist()}")
    env = HumanoidTrackingEnv(cfg, False)
    try:
        for seq, v in m_data.items():
            path = v["path_0"]                               # the path with the smallest total cost
            states, costs = replay(env, path)
            d = np.abs(states - path["simulated_state"]).max(axis=1)
            print(f"seq {seq}: {len(costs)} windows, total cost {path['total_cost']:.4f} (replayed {costs.sum():.4f}); "
                  f"max |replayed - stored state| per window: max {d.max():.3e} ({'bit-exact' if d.max() == 0 else 'differs'})")
            for i, (c, c0) in enumerate(zip(costs, np.asarray(path["cost"]).reshape(-1))):
                print(f"  window {i + 1:3d}: cost {c0:9.4f}  replayed {c:9.4f}  root z {states[i, 2]:.3f} (target {path['target_state'][i][2]:.3f})")
    finally:
        env.close()

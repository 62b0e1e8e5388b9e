"""Sanity probe: a SamCon solve whose reference motion is standing still (and the synthetic walk), reporting the elite
cost per window and the root height of the best path -- the controller must hold a stand indefinitely."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from samcon_b200.config import Config
from samcon_b200.motion import AmassMotionData, make_synthetic_motion
from samcon_b200.samcon import SamCon

kind = sys.argv[1] if len(sys.argv) > 1 else "stand"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
cfg = Config("walk")
m = make_synthetic_motion("walk", T=91, seed=0)
(name, d), = m.items()
if kind == "stand":
    d["pose"][:] = d["pose"][0]          # freeze the first frame
    # feet side by side: zero the hip / knee / ankle swing so that both feet are flat
    d["pose"][:, 7:7 + 4 * 6] = np.tile([0, 0, 0, 1.0], 6)
    from samcon_b200.motion import _lowest_point
    from samcon_b200.urdf import load_links
    d["pose"][:, 2] = -_lowest_point(load_links(None), d["pose"][0]) + 0.001
cfg.motion_seq = name
ld = AmassMotionData(m, name)
sc = SamCon(cfg, ld, 1, 30, S, S // 10)
out, info = sc.sample(save=False)
p = out[name]["path_0"]
print(kind, "cost per window:", np.round(p["cost"].reshape(-1), 2))
print("root height:", np.round(p["simulated_state"][:, 2], 3))
print("target root height:", np.round(p["target_state"][:, 2], 3))
sc.close()

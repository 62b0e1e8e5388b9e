"""Regenerates tests/golden/*.npz.  Needs /root/reference (runs the reference's own SamCon.get_batch_action with
pybullet / pybullet_data / treelib stubbed in sys.modules -- the function itself only uses numpy + scipy).
Run in the build container:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def reference_samcon_class():
    for m in ("pybullet", "pybullet_data", "treelib"):
        sys.modules.setdefault(m, types.ModuleType(m))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    if not hasattr(np, "int"):
        np.int = int
    from samcon.core.samcon import SamCon
    return SamCon


def main():
    sys.path.insert(0, ROOT)
    import yaml
    SamCon = reference_samcon_class()
    out = {}
    for cfg_id, noise_type, seed, n in (("walk", "uniform", 1, 64), ("cartwheel", "uniform", 1, 40), ("walk", "gaussian", 7, 32)):
        y = yaml.safe_load(open(os.path.join(REF, "samcon", "cfg", f"{cfg_id}.yml")))
        cfg = types.SimpleNamespace(noise_type=noise_type, values=np.array([r[1] for r in y["joint_noises"]], dtype=np.float64))
        rng = np.random.default_rng(123)
        state = np.zeros(132)
        state[:3] = [0.0, 0.0, 0.9]
        for k in range(18):
            q = rng.normal(size=4) * np.array([0.5, 0.5, 0.5, 1.0])
            state[3 + 4 * k:7 + 4 * k] = q / np.linalg.norm(q)
        state[75:] = rng.normal(size=57)
        sc = SamCon.__new__(SamCon)
        sc._cfg = cfg
        sc._nSample = n
        np.random.seed(seed)
        a1 = sc.get_batch_action(state)
        a2 = sc.get_batch_action(state)          # second window: RNG stream continues
        key = f"{cfg_id}_{noise_type}"
        out[key + "_state"] = state
        out[key + "_seed"] = np.array(seed)
        out[key + "_values"] = cfg.values
        out[key + "_a1"] = a1
        out[key + "_a2"] = a2
    np.savez_compressed(os.path.join(HERE, "get_batch_action.npz"), **out)
    print("wrote", os.path.join(HERE, "get_batch_action.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

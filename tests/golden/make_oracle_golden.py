"""Regression fixtures of the CPU oracle itself (NOT reference goldens -- the physics of this path is parity-unpinned,
see DESIGN.md §6): a few seeded windows (airborne, standing on the feet, link-vs-link contact, cartwheel hand stand)
with the oracle's outputs, so that an accidental change of the restatement shows up as a diff of committed numbers.

    python tests/golden/make_oracle_golden.py        # rewrites tests/golden/oracle_windows.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def scenarios():
    from helpers import legs_crossed, near_action, rand_state, stand
    from samcon_b200.config import Config
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    rng = np.random.default_rng(2026)
    out = {}
    p = np.array([rand_state(rng, 0.3, 3.0) for _ in range(3)])
    out["air"] = ("walk", p, np.array([near_action(s, rng, 0.1) for s in p]), rand_state(rng, 0.5, 1.0), 60)
    p = np.array([stand(rng, 0.05, 0.2, 0.985 + 0.005 * i) for i in range(3)])
    out["stand"] = ("walk", p, np.array([near_action(s, rng, 0.1) for s in p]), stand(rng), 100)
    p = np.array([legs_crossed(rng, 0.04 + 0.02 * i, 2.0 if i else 0.995, 0.1) for i in range(3)])
    out["self"] = ("walk", p, np.array([legs_crossed(rng, 0.3 + 0.03 * i)[7:75] for i in range(3)]), stand(rng), 100)
    m = make_synthetic_motion("cartwheel", T=76, seed=0)
    (name, _), = m.items()
    ld = AmassMotionData(m, name)
    p = np.repeat(ld.getSpecTimeState(1.2)[None], 2, axis=0)
    tgt = ld.getSpecTimeState(1.3)
    out["cartwheel"] = ("cartwheel", p, np.array([near_action(tgt, rng, 0.05) for _ in range(2)]), tgt, 100)
    return out


def run(sc):
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    res = {}
    for key, (cfgname, parents, actions, target, nsteps) in sc.items():
        o = Oracle(Character.from_config(Config(cfgname)))
        st, cost, info = o.window(parents, actions, target, nsteps)
        res[key] = (st, cost, info)
    return res


if __name__ == "__main__":
    sc = scenarios()
    res = run(sc)
    arrays = {}
    for key in sc:
        cfgname, parents, actions, target, nsteps = sc[key]
        arrays[f"{key}_parents"], arrays[f"{key}_actions"], arrays[f"{key}_target"] = parents, actions, target
        arrays[f"{key}_nsteps"] = np.array(nsteps)
        arrays[f"{key}_state"], arrays[f"{key}_cost"], arrays[f"{key}_info"] = res[key]
    np.savez_compressed(os.path.join(HERE, "oracle_windows.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items() if k.endswith("_state")})

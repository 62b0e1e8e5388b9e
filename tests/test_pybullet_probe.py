"""Hook for the day a real PyBullet is importable (it is not in the build image: pybullet==3.2.5 is the reference's
un-vendored physics).  If `import pybullet` works AND the reference checkout is present, one env.step
of the reference's own HumanoidTrackingEnv is compared with the CPU oracle and the numbers are printed; until then
this skips.  It is a strict test: the day PyBullet is importable, a deviation beyond BASELINE.json's tolerance fails."""
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pybullet_windows.npz")


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="no reference goldens: tests/golden/make_pybullet_golden.py needs pybullet==3.2.5")
def test_committed_pybullet_goldens():
    """Oracle vs windows produced by the REFERENCE env under real PyBullet (tests/golden/make_pybullet_golden.py), at
    BASELINE.json's tolerance.  The file cannot be produced in the build image; once somebody commits it this test is what
    pins the oracle's physics (and lifts "parity unpinned")."""
    from helpers import joint_rms
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    g = np.load(GOLDEN)
    keys = sorted(k[:-6] for k in g.files if k.endswith("_state"))
    assert keys
    for key in keys:
        cfgname = "cartwheel" if key == "cartwheel" else "walk"
        o = Oracle(Character.from_config(Config(cfgname)))
        parents, actions, target = g[f"{key}_parents"], g[f"{key}_actions"], g[f"{key}_target"]
        for p, a, rs, rc in zip(parents, actions, g[f"{key}_state"], g[f"{key}_cost"]):
            st, cost, _ = o.window(p[None], a[None], target, 100)
            assert joint_rms(st, rs[None])[0] < 1e-3 and abs(cost[0] - rc) <= 1e-3 * abs(rc), key




@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout absent")
def test_one_window_against_real_pybullet(oracle, walk_cfg):
    pytest.importorskip("pybullet")
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from helpers import near_action, stand
    from samcon.utils.config import Config as RefConfig          # the reference's own classes, unmodified
    from envs.humanoid_tracking_env import HumanoidTrackingEnv as RefEnv
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        env = RefEnv(RefConfig("walk", base_dir=""), False)
        rng = np.random.default_rng(0)
        s0, tgt = stand(rng, 0.03, 0.1, 0.99), stand(rng)
        act = near_action(s0, rng, 0.1)
        env._sim_agent.set_state(s0)
        env._kin_agent.set_state(tgt)
        ref_state, ref_cost, ref_info = env.step(act)
    finally:
        os.chdir(cwd)
    st, cost, info = oracle.window(s0[None], act[None], tgt, walk_cfg.fps_sim // walk_cfg.fps_act)
    from helpers import joint_rms
    rms = joint_rms(st, np.asarray(ref_state)[None])[0]
    print(f"oracle vs PyBullet: joint rms {rms:.3e} rad, cost {cost[0]:.6f} vs {ref_cost:.6f}")
    assert rms < 1e-3 and abs(cost[0] - ref_cost) <= 1e-3 * abs(ref_cost)

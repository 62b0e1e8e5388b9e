"""Generator of REFERENCE goldens for the physics of this path -- to be run by somebody who has pybullet==3.2.5
(requirements.txt:5 of the reference), which the build image has not (no network): it runs the reference's own
HumanoidTrackingEnv on a few seeded (state, action, target) triples and stores what env.step returns.

    cd /root/reference && python /root/repo/tests/golden/make_pybullet_golden.py      # writes tests/golden/pybullet_windows.npz

tests/test_pybullet_probe.py::test_committed_pybullet_goldens compares the oracle with that file whenever it exists
(and is skipped until then).  The scenarios are the ones of tests/golden/make_oracle_golden.py (airborne, standing,
link-vs-link contact), so the file also says which of the [BULLET] assumptions of DESIGN.md §2 hold."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get("SAMCON_REFERENCE", "/root/reference")


def main():
    import pybullet  # noqa: F401  -- fails here if the wheel is absent, which is the point
    from make_oracle_golden import scenarios
    sys.path.insert(0, REF)
    from envs.humanoid_tracking_env import HumanoidTrackingEnv as RefEnv
    from samcon.utils.config import Config as RefConfig
    cwd = os.getcwd()
    os.chdir(REF)
    out = {"pybullet_version": np.array(getattr(pybullet, "__version__", "unknown"))}
    try:
        for key, (cfgname, parents, actions, target, nsteps) in scenarios().items():
            if nsteps != 100:
                continue                                    # env.step is one whole window (fps_sim // fps_act steps)
            env = RefEnv(RefConfig(cfgname, base_dir=""), False)
            st, cost, info = [], [], []
            for p, a in zip(parents, actions):
                env._sim_agent.set_state(p)
                env._kin_agent.set_state(target)
                s, c, i = env.step(a)
                st.append(s); cost.append(c); info.append(i)
            out[f"{key}_parents"], out[f"{key}_actions"], out[f"{key}_target"] = parents, actions, target
            out[f"{key}_state"], out[f"{key}_cost"], out[f"{key}_info"] = np.array(st), np.array(cost), np.array(info)
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "pybullet_windows.npz"), **out)
    print("wrote", os.path.join(HERE, "pybullet_windows.npz"))


if __name__ == "__main__":
    main()

"""Why two GPU parity tests apply BASELINE.json's tolerance to the distribution instead of to every sample
(tests/test_gpu_rollout.py::test_walk_window_and_elites: >= 95 % inside; tests/test_gpu_configs.py cartwheel: >= 75 %).

The kernel source compiled for the host twice -- once as it runs on the device (float32, tests/emu/emu.cpp) and once with
every float turned into a double (tests/emu/emu64.cpp) -- against the float64 oracle on those two windows:

* the double build agrees with the oracle to ~1e-6 rad on the well-conditioned samples: the kernel's algorithms
  (articulated-body recursions, contact-space Gauss-Seidel) compute the same physics as the oracle's dense ones;
* the samples on which the float32 build leaves the tolerance are chaotic ones -- feet or hands chattering on the ground,
  where a perturbation grows by orders of magnitude inside the window: there the double build deviates from the oracle
  too (its 1e-16 differences reach 1e-6 .. 1e-3), and far more than on the rest.  Rounding is amplified by the dynamics,
  not produced by the kernel.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import joint_rms, stand
from samcon_b200.abi import make_model

HERE = os.path.dirname(os.path.abspath(__file__))


def _build(name):
    src = os.path.join(HERE, "emu", name + ".cpp")
    lib = os.path.join(HERE, "emu", "libsc_" + name + ".so")
    deps = [src] + [os.path.join(HERE, "..", "samcon_b200", "csrc", f) for f in ("sc_core.cuh", "sc_tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", lib, src])
    return C.CDLL(lib)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _run32(lib, character, parents, actions, target, nsteps, children):
    m, keep = make_model(character)
    P, A, T = (np.ascontiguousarray(x, np.float32) for x in (parents, actions, target))
    S = len(A)
    st, cost, info = np.zeros((S, 132), np.float32), np.zeros(S, np.float32), np.zeros((S, 5), np.float32)
    err = C.create_string_buffer(256)
    assert lib.emu_window(C.byref(m), _p(P), len(P), None, children, _p(A), _p(T), 1, None, S, nsteps, _p(st), _p(cost), _p(info),
                          None, None, err, 256) == 0, err.value
    return st.astype(np.float64), cost.astype(np.float64)


def _run64(lib, character, parents, actions, target, nsteps, children):
    m, keep = make_model(character)
    P, A, T = (np.ascontiguousarray(x, np.float64) for x in (parents, actions, target))
    S = len(A)
    st, cost, info = np.zeros((S, 132)), np.zeros(S), np.zeros((S, 5))
    err = C.create_string_buffer(256)
    assert lib.emu64_window(C.byref(m), _p(P), len(P), children, _p(A), _p(T), 1, S, nsteps, _p(st), _p(cost), _p(info), err, 256) == 0, err.value
    return st, cost


def _report(tag, r32, r64):
    bad = r32 > 1e-3
    print(f"{tag}: float32 build vs oracle: median {np.median(r32):.1e}, {bad.sum()} of {len(r32)} samples above 1e-3 rad (max {r32.max():.1e}); "
          f"double build vs oracle: median {np.median(r64):.1e}, on those samples median {np.median(r64[bad]) if bad.any() else 0:.1e}, "
          f"on the others max {r64[~bad].max():.1e}")
    return bad


def test_walk_noise_window_outliers_are_chaotic_samples(character, oracle, walk_cfg):
    """The window of test_walk_window_and_elites, widened to 240 samples: cfg noise (0.8 rad on the hips) on standing parents."""
    from samcon_b200.sampling import get_batch_action
    e32, e64 = _build("emu"), _build("emu64")
    rng = np.random.default_rng(4)
    E, children = 24, 10
    parents = np.array([stand(rng, 0.03, 0.1, 0.985) for _ in range(E)])
    target = stand(rng, 0.03, 0.1, 0.985)
    actions = get_batch_action(target, E * children, walk_cfg.values, "uniform", np.random.RandomState(1))
    ost, ocost, _ = oracle.window(parents, actions, target, 100)
    s32, c32 = _run32(e32, character, parents, actions, target, 100, children)
    s64, c64 = _run64(e64, character, parents.astype(np.float32), actions.astype(np.float32), target.astype(np.float32), 100, children)
    # (the double build gets the float32-rounded inputs the device gets; the oracle the float64 ones, like the GPU tests)
    r32, r64 = joint_rms(s32, ost), joint_rms(s64, ost)
    bad = _report("walk.yml noise window", r32, r64)
    assert np.median(r64) < 1e-5 and np.median(r32) < 5e-5
    assert (r32 < 1e-3).mean() >= 0.95                                    # the bar the GPU test applies
    if bad.any():
        # the float32 outliers are the samples the dynamics amplify: the double build's own deviation there is orders of
        # magnitude above its deviation elsewhere
        assert np.median(r64[bad]) > 20 * np.median(r64[~bad])


@pytest.mark.parametrize("t0", [0.9, 1.2])
def test_cartwheel_window_outliers_are_chaotic_samples(t0):
    """The windows of test_cartwheel_window_parity (cartwheel.yml): hands striking the ground (t0 = 0.9 s of the synthetic
    cartwheel, where the GPU test accepts 75 % of the samples inside the tolerance) and the hand stand (t0 = 1.2 s, 95 %)."""
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    from samcon_b200.sampling import get_batch_action
    e32, e64 = _build("emu"), _build("emu64")
    cfg = Config("cartwheel")
    ch = Character.from_config(cfg)
    o = Oracle(ch)
    m = make_synthetic_motion("cartwheel", T=76, seed=0)
    (name, _), = m.items()
    ld = AmassMotionData(m, name)
    start, target = ld.getSpecTimeState(t0), ld.getSpecTimeState(t0 + 0.1)
    E, children = 12, 4
    parents = np.repeat(start[None], E, axis=0)
    actions = get_batch_action(target, E * children, 0.25 * cfg.values, "uniform", np.random.RandomState(3))
    ost, ocost, _ = o.window(parents, actions, target, 100)
    s32, _ = _run32(e32, ch, parents, actions, target, 100, children)
    s64, _ = _run64(e64, ch, parents.astype(np.float32), actions.astype(np.float32), target.astype(np.float32), 100, children)
    r32, r64 = joint_rms(s32, ost), joint_rms(s64, ost)
    bad = _report(f"cartwheel t0 = {t0}", r32, r64)
    assert np.median(r32) < 1e-4 and (r32 < 1e-3).mean() >= (0.75 if t0 == 0.9 else 0.95)       # the bars the GPU test applies
    if bad.any():
        assert np.median(r64[bad]) > 20 * np.median(r64[~bad])

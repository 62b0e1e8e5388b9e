"""The rollout kernel's arithmetic (samcon_b200/csrc/sc_core.cuh compiled for the host as a lane loop, float32)
against the float64 oracle, so kernel logic regressions are caught on a box without a GPU.  The emulation is a
test aid only -- the package never loads it; the GPU parity proper is tests/test_gpu_rollout.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import joint_rms, near_action, rand_state, stand
from samcon_b200.abi import make_model

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu.cpp")
    lib = os.path.join(HERE, "emu", "libsc_emu.so")
    deps = [src] + [os.path.join(HERE, "..", "samcon_b200", "csrc", f) for f in ("sc_core.cuh", "sc_tables.h")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wno-unknown-pragmas", "-o", lib, src])
    return C.CDLL(lib)


def _window(emu, character, parents, actions, targets, nsteps, children=1, parent_idx=None, sample_seq=None,
            parent_cache=None, return_cache=False):
    m, keep = make_model(character)
    parents = np.ascontiguousarray(parents, np.float32)
    actions = np.ascontiguousarray(actions, np.float32)
    targets = np.ascontiguousarray(targets, np.float32).reshape(-1, 132)
    S, E = len(actions), len(parents)
    st = np.zeros((S, 132), np.float32)
    cost = np.zeros(S, np.float32)
    info = np.zeros((S, 5), np.float32)
    err = C.create_string_buffer(256)

    def p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None
    pi = np.ascontiguousarray(parent_idx, np.int32) if parent_idx is not None else None
    sq = np.ascontiguousarray(sample_seq, np.int32) if sample_seq is not None else None
    pc = np.ascontiguousarray(parent_cache, np.float32) if parent_cache is not None else None
    oc = np.zeros((S, 16), np.float32) if return_cache else None
    rc = emu.emu_window(C.byref(m), p(parents), E, p(pi), children, p(actions), p(targets), len(targets), p(sq), S, nsteps,
                        p(st), p(cost), p(info), p(pc), p(oc), err, 256)
    assert rc == 0, err.value
    return (st, cost, info, oc) if return_cache else (st, cost, info)


def test_airborne(emu, character, oracle):
    rng = np.random.default_rng(2)
    parents = np.array([rand_state(rng, 0.3, 3.0) for _ in range(10)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = rand_state(rng, 0.5, 1.0)
    st, cost, info = _window(emu, character, parents, actions, target, 50)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 50)
    assert joint_rms(st, ost).max() < 1e-4
    np.testing.assert_allclose(cost, ocost, rtol=1e-3)


def test_contact_window(emu, character, oracle):
    rng = np.random.default_rng(3)
    parents = np.array([stand(rng, 0.05, 0.2, 0.98 + 0.005 * i) for i in range(9)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = stand(rng)
    st, cost, info = _window(emu, character, parents, actions, target, 100)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    assert joint_rms(st, ost).max() < 1e-3
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=1e-4)
    np.testing.assert_allclose(cost, ocost, rtol=1e-3)
    np.testing.assert_allclose(info, oinfo, rtol=5e-3, atol=1e-5)


def test_parent_mapping_and_sequences(emu, character, oracle):
    """children = S/E mapping (sample k <- elite k // children), explicit parent_idx, per-sample target sequences."""
    rng = np.random.default_rng(4)
    parents = np.array([stand(rng, 0.05, 0.2, 1.0) for _ in range(3)])
    targets = np.array([stand(rng) for _ in range(2)])
    actions = np.array([near_action(parents[k // 2], rng, 0.1) for k in range(6)])
    st, cost, _ = _window(emu, character, parents, actions, targets[0], 10, children=2)
    ost, ocost, _ = oracle.window(parents, actions, targets[0], 10)
    assert joint_rms(st, ost).max() < 1e-4
    pidx = np.array([2, 0, 1, 1, 0, 2, 2])
    seq = np.array([0, 1, 1, 0, 1, 0, 1])
    actions = np.array([near_action(parents[p], rng, 0.1) for p in pidx])
    st, cost, _ = _window(emu, character, parents, actions, targets, 10, parent_idx=pidx, sample_seq=seq)
    for k in range(7):
        os_, oc, _ = oracle.window(parents[pidx[k]][None], actions[k][None], targets[seq[k]], 10)
        assert joint_rms(st[k][None], os_).max() < 1e-4
        np.testing.assert_allclose(cost[k], oc[0], rtol=1e-3)


def test_self_contact_window(emu, character, oracle):
    """Link-vs-link contacts (thighs / shins / feet squeezed together by the PD targets, airborne and standing):
    two-body contact rows, per-contact friction, general contact frames."""
    from helpers import legs_crossed
    rng = np.random.default_rng(7)
    parents = np.array([legs_crossed(rng, 0.03 + 0.01 * i, 2.0 if i % 2 else 0.995, 0.1) for i in range(8)])
    actions = np.array([legs_crossed(rng, 0.3 + 0.02 * i)[7:75] for i in range(8)])
    target = stand(rng)
    n_self = sum((oracle.contacts_full(oracle.step(p, a, 60))[:, 1] >= 0).sum() for p, a in zip(parents, actions))
    assert n_self >= 6                                              # the scenario really produces link-link contacts
    st, cost, info = _window(emu, character, parents, actions, target, 100)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    assert joint_rms(st, ost).max() < 1e-3
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=2e-4)
    np.testing.assert_allclose(cost, ocost, rtol=1e-3)


def test_self_collision_switch(emu, walk_cfg):
    """cfg self_collision: False (npair = 0) must reproduce the oracle built without pairs."""
    from helpers import legs_crossed
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    ch = Character.from_config(walk_cfg, SimParams(self_collision=False))
    o = Oracle(ch)
    rng = np.random.default_rng(8)
    parents = np.array([legs_crossed(rng, 0.05, 2.0, 0.1) for _ in range(4)])
    actions = np.array([legs_crossed(rng, 0.35)[7:75] for _ in range(4)])
    target = stand(rng)
    st, cost, _ = _window(emu, ch, parents, actions, target, 300)
    ost, ocost, _ = o.window(parents, actions, target, 300)
    assert joint_rms(st, ost).max() < 1e-4
    assert (2 * np.arccos(np.abs(st[:, 10]))).min() > 0.3          # the legs passed through each other


def test_contact_cap_paths(emu, character, oracle):
    """More candidates inside the margin than max_contacts: the parallel pick (<= 32 hits: both feet flat plus
    link-vs-link hits) and the serial fallback (a humanoid lying on the ground: > 32 hits) must keep the same
    deepest-8 set as the oracle."""
    from helpers import legs_crossed
    rng = np.random.default_rng(9)
    flat = []
    for i in range(3):                                              # lying on the back / side, a few cm above ground
        s = stand(rng, 0.02, 0.1, 0.09 + 0.01 * i)
        s[3:7] = [0, 0, 0, 1] if i < 2 else [0, np.sin(0.4), 0, np.cos(0.4)]
        flat.append(s)
    feet = [legs_crossed(rng, 0.12 + 0.02 * i, 0.97, 0.05) for i in range(3)]   # feet 1 cm into the ground, shins touching
    parents = np.array(flat + feet)
    counts = [len(oracle.contacts_full(p)) for p in parents]
    assert counts == [character.sim.max_contacts] * 6
    actions = np.array([near_action(p, rng, 0.05) for p in parents])
    target = stand(rng)
    st, cost, _ = _window(emu, character, parents, actions, target, 6)
    ost, ocost, _ = oracle.window(parents, actions, target, 6)
    assert joint_rms(st, ost).max() < 2e-4
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=1e-4)


def test_contact_cache_chain(emu, character, oracle):
    """The contact cache is part of a sample's state: a window cut in two and resumed from (state, cache) gives the
    bits of the uncut window (that is what restoring a .bullet snapshot does for the reference,
    humanoid_tracking_env.py:178-182); resumed from the state alone it does not.  The cache rows agree with the oracle's."""
    rng = np.random.default_rng(11)
    parents = np.array([stand(rng, 0.05, 0.2, 0.98 + 0.005 * i) for i in range(6)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = stand(rng)
    full, cfull, _, cache_full = _window(emu, character, parents, actions, target, 100, return_cache=True)
    half, _, _, cache_half = _window(emu, character, parents, actions, target, 50, return_cache=True)
    assert (cache_half[:, 0] >= 0).sum() >= 4 and (cache_half[:, 8:] >= 0).all() and cache_half[:, 8:].max() > 0.01
    rest, crest, _, cache_rest = _window(emu, character, half, actions, target, 50, parent_cache=cache_half, return_cache=True)
    np.testing.assert_array_equal(rest, full)
    np.testing.assert_array_equal(crest, cfull)
    np.testing.assert_array_equal(cache_rest, cache_full)
    cold, _, _ = _window(emu, character, half, actions, target, 50)
    assert np.abs(cold - full).max() > 0                                          # the cache matters
    # against the oracle: same candidate ids in the same slots, impulses to float accuracy; and the same chain
    ost, ocost, _, ocache = oracle.window(parents, actions, target, 50, return_cache=True)
    np.testing.assert_array_equal(cache_half[:, :8], ocache[:, :8])
    # (how a foot's load splits over its corners is ill-conditioned; the total is not)
    np.testing.assert_allclose(cache_half[:, 8:], ocache[:, 8:], rtol=0.1, atol=1e-3)
    np.testing.assert_allclose(cache_half[:, 8:].sum(1), ocache[:, 8:].sum(1), rtol=5e-2, atol=1e-4)      # one sub-step's impulse of a settling contact is noisy
    ost2, ocost2, _ = oracle.window(ost, actions, target, 50, parent_cache=ocache)
    ofull, ocfull, _ = oracle.window(parents, actions, target, 100)
    np.testing.assert_array_equal(ost2, ofull)
    assert joint_rms(rest, ost2).max() < 1e-3


def test_warmstart_off_matches_stateless_oracle(emu, walk_cfg):
    """warmstart = 0 is the round-1 stateless contact model, on both sides."""
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    ch = Character.from_config(walk_cfg, SimParams(warmstart=0.0))
    o = Oracle(ch)
    rng = np.random.default_rng(12)
    parents = np.array([stand(rng, 0.05, 0.2, 0.98 + 0.005 * i) for i in range(4)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = stand(rng)
    st, cost, _, cache = _window(emu, ch, parents, actions, target, 60, return_cache=True)
    ost, ocost, _ = o.window(parents, actions, target, 60)
    assert joint_rms(st, ost).max() < 1e-3
    np.testing.assert_allclose(cost, ocost, rtol=1e-3)
    again, _, _ = _window(emu, ch, st, actions, target, 20, parent_cache=cache)
    cold, _, _ = _window(emu, ch, st, actions, target, 20)
    np.testing.assert_array_equal(again, cold)                                      # without warm start the cache is inert


def _box_vs_round_poses(oracle, n, seed=5):
    """Seeded search for airborne poses (random left leg and arms) in which a box link -- a foot, an upper arm, a forearm --
    is within the contact margin of a capsule / sphere link that is not its ancestor (found: the left foot's box against the
    right shin's capsule)."""
    from helpers import rand_quat
    rng = np.random.default_rng(seed)
    boxes = {3, 6, 13, 14, 17, 18}                           # URDF link indices of the box links
    out = []
    while len(out) < n:
        s = stand(rng, 0.0, 0.0, 2.0)
        s[0:2] = 0
        for j in (0, 1, 2, 12, 13, 15, 16):
            s[7 + 4 * j:11 + 4 * j] = rand_quat(rng, 1.2)
        c = oracle.contacts_full(s)
        c = c[c[:, 1] >= 0]
        if any((int(r[0]) in boxes) != (int(r[1]) in boxes) and -0.02 < r[11] < 0.015 for r in c):
            out.append(s)
    return np.array(out)


def test_box_contacts_exact_narrow_phase(emu, walk_cfg):
    """Link-vs-link pairs with a box (SimParams.exact_box_contacts=True): a foot box against the other leg's shin capsule (box
    vs swept segment, exact) and feet against each other (box vs box: the second as its capsule) -- the kernel's float32
    safeguarded Newton against the oracle's float64 one."""
    from helpers import legs_crossed
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    character = Character.from_config(walk_cfg, SimParams(exact_box_contacts=True))
    oracle = Oracle(character)
    rng = np.random.default_rng(13)
    parents = np.concatenate([_box_vs_round_poses(oracle, 6), np.array([legs_crossed(rng, 0.10 + 0.02 * i, 2.0, 0.05) for i in range(2)])])
    # PD targets = the poses themselves, slightly perturbed: the touching links keep pressing / sliding on each other
    actions = np.array([near_action(p, rng, 0.05) for p in parents])
    target = stand(rng)
    boxes = {3, 6, 13, 14, 17, 18}
    seen = set()
    for p, a in zip(parents, actions):
        for t in (0, 20, 50):
            c = oracle.contacts_full(oracle.step(p, a, t) if t else p)
            for r in c[c[:, 1] >= 0]:
                la, lb = int(r[0]), int(r[1])
                seen.add("box-vs-box" if la in boxes and lb in boxes else "box-vs-round" if (la in boxes) != (lb in boxes) else "round")
    assert {"box-vs-round", "box-vs-box"} <= seen, seen
    st, cost, info = _window(emu, character, parents, actions, target, 100)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    rms = joint_rms(st, ost)
    print("box contact windows: joint rms", rms)
    assert rms.max() < 1e-3
    np.testing.assert_allclose(cost, ocost, rtol=1e-3)
    # the capsule stand-in (the default) is a different model: the switch changes these windows
    o1 = Oracle(Character.from_config(walk_cfg, SimParams(exact_box_contacts=False)))
    ost1, _, _ = o1.window(parents, actions, target, 100)
    assert joint_rms(ost1, ost).max() > 1e-3

"""GPU parity: the CUDA rollout (through the C ABI) against the float64 oracle on the same seeded inputs.

Tolerances are BASELINE.json's: per-sample joint-state RMS error <= 1e-3 rad and cost within 1e-3 relative
after one window; elite index sets equal whenever costs are separated by more than that tolerance.

Contact-rich samples with the cfg's large hip noise (0.8 rad) are occasionally *chaotic*: rounding-level
differences grow ~50x per simulated millisecond while feet chatter on the ground.  The float64 build of the very
same kernel source deviates from the oracle on exactly those samples (tests/test_f64_kernel_build.py shows it on this
window), so for the cfg-noise window the bar is applied to the distribution: >= 98 % of the samples inside the
tolerance, tight median, bounded worst case.
"""
import os

import numpy as np
import pytest

from helpers import joint_rms, near_action, rand_state, stand

pytestmark = pytest.mark.gpu

JOINT_TOL = 1e-3
COST_RTOL = 1e-3


def _run(pool, parents, actions, target, nsteps, children=None):
    st, cost, info = pool.window(np.asarray(parents, np.float32), np.asarray(actions, np.float32),
                                 np.asarray(target, np.float32), nsteps, children=children)
    return st.cpu().numpy(), cost.cpu().numpy(), info.cpu().numpy()


def test_cost_only(pool, oracle):
    rng = np.random.default_rng(1)
    parents = np.array([rand_state(rng, 1.0, 1.2) for _ in range(37)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = rand_state(rng, 0.5, 1.0)
    st, cost, info = _run(pool, parents, actions, target, 0, 1)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 0)
    np.testing.assert_allclose(st, ost, atol=2e-6)
    np.testing.assert_allclose(cost, ocost, rtol=1e-4)
    np.testing.assert_allclose(info, oinfo, rtol=1e-4, atol=1e-6)


def test_airborne_window(pool, oracle):
    rng = np.random.default_rng(2)
    parents = np.array([rand_state(rng, 0.3, 3.0) for _ in range(24)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = rand_state(rng, 0.5, 1.0)
    st, cost, info = _run(pool, parents, actions, target, 100, 1)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    assert joint_rms(st, ost).max() < JOINT_TOL
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=1e-4)
    np.testing.assert_allclose(cost, ocost, rtol=COST_RTOL)


def test_contact_window(pool, oracle):
    """Standing / landing humanoids: foot boxes, partial support, speculative contacts."""
    rng = np.random.default_rng(3)
    parents = np.array([stand(rng, 0.05, 0.2, 0.98 + 0.004 * i) for i in range(16)])
    actions = np.array([near_action(p, rng, 0.1) for p in parents])
    target = stand(rng)
    st, cost, info = _run(pool, parents, actions, target, 100, 1)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    assert joint_rms(st, ost).max() < JOINT_TOL
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=1e-4)
    np.testing.assert_allclose(cost, ocost, rtol=COST_RTOL)
    np.testing.assert_allclose(info, oinfo, rtol=5e-3, atol=1e-5)


def test_walk_window_and_elites(pool, oracle):
    """walk.yml-shaped window, reduced: E parents x 10 children, cfg noise, elite set vs oracle."""
    from samcon_b200.sampling import get_batch_action
    rng = np.random.default_rng(4)
    E, children = 12, 10
    parents = np.array([stand(rng, 0.03, 0.1, 0.985) for _ in range(E)])
    target = stand(rng, 0.03, 0.1, 0.985)
    from samcon_b200.config import Config
    cfg = Config("walk")
    actions = get_batch_action(target, E * children, cfg.values, "uniform", np.random.RandomState(1))
    st, cost, info = _run(pool, parents, actions, target, 100)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    rms = joint_rms(st, ost)
    crel = np.abs(cost - ocost) / np.abs(ocost)
    print(f"joint rms: median {np.median(rms):.2e} p95 {np.quantile(rms, .95):.2e} max {rms.max():.2e}; "
          f"cost rel: median {np.median(crel):.2e} max {crel.max():.2e}")
    assert np.median(rms) < 5e-5 and (rms < JOINT_TOL).mean() >= 0.98 and rms.max() < 0.05
    assert np.median(crel) < 1e-4 and (crel < COST_RTOL).mean() >= 0.98
    idx, ec = pool.select(cost, E)
    gi = idx.cpu().numpy()[0]
    oi = oracle.topk(ocost, E)
    # identical wherever the oracle's costs are separated by more than the tolerance
    srt = np.sort(ocost)
    sep = np.diff(srt)[:E] > 2 * COST_RTOL * srt[1:E + 1]
    if sep.all():
        assert list(gi) == list(oi)
    else:
        assert set(gi[: np.argmin(sep)]) == set(oi[: np.argmin(sep)])
    assert list(gi) == list(np.lexsort((np.arange(len(cost)), cost))[:E])


def test_step_and_cost_entry_points(pool, oracle):
    rng = np.random.default_rng(5)
    states = np.array([stand(rng, 0.05, 0.2, 1.0) for _ in range(9)])
    actions = np.array([near_action(p, rng, 0.1) for p in states])
    import torch
    d = torch.tensor(states, dtype=torch.float32, device="cuda")
    pool.step(d, actions, 20)
    ost = np.array([oracle.step(s, a, 20) for s, a in zip(states, actions)])
    assert joint_rms(d.cpu().numpy(), ost).max() < JOINT_TOL
    kin = np.array([stand(rng) for _ in range(9)])
    c = pool.cost(ost.astype(np.float32), kin.astype(np.float32)).cpu().numpy()
    for i in range(9):
        tot, terms = oracle.cost(ost[i], kin[i])
        np.testing.assert_allclose(c[i, 0], tot, rtol=1e-4)
        np.testing.assert_allclose(c[i, 1:], terms, rtol=1e-4, atol=1e-7)


def test_ragged_sizes(pool, oracle):
    """S not a multiple of the 8-sample tile, S=1, explicit parent_idx."""
    rng = np.random.default_rng(6)
    parents = np.array([stand(rng, 0.05, 0.2, 1.0) for _ in range(3)])
    target = stand(rng)
    for S in (1, 7, 13):
        pidx = rng.integers(0, 3, size=S).astype(np.int32)
        actions = np.array([near_action(parents[p], rng, 0.1) for p in pidx])
        st, cost, info = pool.window(parents.astype(np.float32), actions.astype(np.float32), target.astype(np.float32), 10,
                                     parent_idx=pidx)
        ost, ocost, _ = oracle.window(parents[pidx], actions, target, 10)
        assert joint_rms(st.cpu().numpy(), ost).max() < JOINT_TOL
        np.testing.assert_allclose(cost.cpu().numpy(), ocost, rtol=COST_RTOL)


def test_self_contact_window(pool, oracle):
    """Link-vs-link contact rows on the GPU: legs squeezed together by the PD targets, airborne and standing."""
    from helpers import legs_crossed
    rng = np.random.default_rng(7)
    n = 24
    parents = np.array([legs_crossed(rng, 0.03 + 0.004 * i, 2.0 if i % 2 else 0.995, 0.1) for i in range(n)])
    actions = np.array([legs_crossed(rng, 0.3 + 0.008 * i)[7:75] for i in range(n)])
    target = stand(rng)
    n_self = sum((oracle.contacts_full(oracle.step(p, a, 60))[:, 1] >= 0).sum() for p, a in zip(parents[:8], actions[:8]))
    assert n_self >= 6
    st, cost, info = _run(pool, parents, actions, target, 100, 1)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    rms = joint_rms(st, ost)
    print(f"self-contact joint rms: median {np.median(rms):.2e} max {rms.max():.2e}")
    assert rms.max() < JOINT_TOL
    np.testing.assert_allclose(st[:, :3], ost[:, :3], atol=2e-4)
    np.testing.assert_allclose(cost, ocost, rtol=COST_RTOL)


def test_committed_fixture_windows(pool):
    """The committed oracle fixtures (tests/golden/oracle_windows.npz) through the C ABI: the GPU path against
    numbers that travel with the repo -- the walk-cfg scenarios on the walk.yml pool, the cartwheel hand stand on a pool
    built from cartwheel.yml (its own cost weights)."""
    import os
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    from samcon_b200.pool import GpuSamplePool
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_windows.npz"))
    for key in ("air", "stand", "self"):
        st, cost, info = _run(pool, g[f"{key}_parents"], g[f"{key}_actions"], g[f"{key}_target"], int(g[f"{key}_nsteps"]), 1)
        assert joint_rms(st, g[f"{key}_state"]).max() < JOINT_TOL
        np.testing.assert_allclose(cost, g[f"{key}_cost"], rtol=COST_RTOL)
    cw = GpuSamplePool(Character.from_config(Config("cartwheel")), 0)
    try:
        key = "cartwheel"
        st, cost, info = _run(cw, g[f"{key}_parents"], g[f"{key}_actions"], g[f"{key}_target"], int(g[f"{key}_nsteps"]), 1)
        rms = joint_rms(st, g[f"{key}_state"])
        print(f"cartwheel fixture: joint rms {rms}, cost {cost} vs {g[f'{key}_cost']}")
        assert rms.max() < JOINT_TOL
        np.testing.assert_allclose(cost, g[f"{key}_cost"], rtol=COST_RTOL)
    finally:
        cw.close()


def test_contact_cache_travels_with_the_state(pool, oracle):
    """Warm-started contacts: the cache row is part of a sample's state.  (i) A window cut in two and resumed from
    (state, cache) is the uncut window bit for bit -- what restoring a .bullet snapshot gives the reference
    (humanoid_tracking_env.py:178-182); resumed from the state alone it is not.  (ii) Cache rows: the oracle's candidate
    ids in the oracle's slots.  (iii) A chain of windows through pool.sim/keep (which carries the caches) agrees with the
    oracle chained the same way."""
    import torch
    rng = np.random.default_rng(21)
    parents = np.array([stand(rng, 0.05, 0.2, 0.98 + 0.004 * i) for i in range(12)]).astype(np.float32)
    actions = np.array([near_action(p, rng, 0.1) for p in parents]).astype(np.float32)
    target = stand(rng).astype(np.float32)
    full = pool.window(parents, actions, target, 100, children=1, out_cache=True)
    half = pool.window(parents, actions, target, 50, children=1, out_cache=True)
    rest = pool.window(half[0], actions, target, 50, children=1, parent_cache=half[3], out_cache=True)
    for a, b in zip(full, rest):
        assert torch.equal(a, b)
    cold = pool.window(half[0], actions, target, 50, children=1)
    assert not torch.equal(cold[0], full[0])
    ost, ocost, _, ocache = oracle.window(parents, actions, target, 50, return_cache=True)
    hc = half[3].cpu().numpy()
    assert (hc[:, 0] >= 0).sum() >= 8
    np.testing.assert_array_equal(hc[:, :8], ocache[:, :8])
    np.testing.assert_allclose(hc[:, 8:].sum(1), ocache[:, 8:].sum(1), rtol=5e-2, atol=1e-4)
    ost2, ocost2, _ = oracle.window(ost, actions, target, 50, parent_cache=ocache)
    assert joint_rms(rest[0].cpu().numpy(), ost2).max() < JOINT_TOL
    np.testing.assert_allclose(rest[1].cpu().numpy(), ocost2, rtol=COST_RTOL)
    # (iii) host protocol: init / sim / keep
    from oracle_pool import OraclePool
    op = OraclePool(oracle)
    E, children = 4, 3
    pool.init(parents[0], E); op.init(parents[0], E)
    acts = np.array([near_action(parents[0], rng, 0.08) for _ in range(E * children)])
    for w in range(3):
        gs, gc, gi = pool.sim(target, acts.astype(np.float32), 40)
        os_, oc, oi = op.sim(target, acts.astype(np.float32), 40)
        assert joint_rms(gs, os_).max() < JOINT_TOL
        np.testing.assert_allclose(gc, oc, rtol=COST_RTOL)
        keep = np.lexsort((np.arange(len(oc)), oc))[:E]
        pool.keep(keep); op.keep(keep)


def test_contact_cap_statistics(pool):
    """sc_contact_stats: how often more candidate points were inside the margin than max_contacts, and how often more
    than that were actually touching.  A humanoid lying on the ground overflows in every sub-step, an airborne one never."""
    rng = np.random.default_rng(22)
    pool.contact_stats(reset=True)
    air = np.array([rand_state(rng, 0.3, 3.0) for _ in range(5)]).astype(np.float32)
    pool.window(air, air[:, 7:75].copy(), air[0], 10, children=1)
    ov, n, hard = pool.contact_stats()
    assert (ov, n, hard) == (0, 5 * 10 * 2, 0)
    lying = stand(rng, 0.02, 0.0, 0.08)
    lying[3:7] = [0, 0, 0, 1]
    lying = np.repeat(lying[None], 3, axis=0).astype(np.float32)
    pool.window(lying, lying[:, 7:75].copy(), lying[0], 10, children=1)
    ov, n, hard = pool.contact_stats()
    assert n == 3 * 10 * 2 and ov >= n // 2 and 0 < hard <= ov
    assert pool.contact_stats() == (0, 0, 0)


def test_entry_points_restore_the_callers_device(character):
    """Every C-ABI call runs on the handle's device and puts the caller's current device back."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from samcon_b200.pool import GpuSamplePool
    other = GpuSamplePool(character, 1)
    try:
        assert torch.cuda.current_device() == 0
        rng = np.random.default_rng(23)
        p = torch.tensor(np.array([stand(rng) for _ in range(2)]), dtype=torch.float32, device="cuda:1")
        other.window(p, p[:, 7:75].contiguous(), p[0], 5, children=1)
        other.select(torch.rand(100, device="cuda:1"), 10)
        assert torch.cuda.current_device() == 0
    finally:
        other.close()


def test_select_edge_cases(pool):
    """Trajectory.select semantics at the edges: ties by index, NaN / inf costs last, a segment shorter than E padded
    with -1 / +inf, E == n, one-element segments."""
    import torch
    c = torch.tensor([3.0, 1.0, float("nan"), 1.0, float("inf"), -2.0, 1.0, float("-inf")], device="cuda")
    idx, ec = pool.select(c, 8)
    assert idx[0].tolist() == [7, 5, 1, 3, 6, 0, 4, 2]
    assert torch.isnan(ec[0, 7]) and ec[0, 6] == float("inf")
    seg = torch.tensor([0, 3, 4, 8], dtype=torch.int32, device="cuda")
    idx, ec = pool.select(c, 4, seg_offsets=seg)
    assert idx.tolist() == [[1, 0, 2, -1], [3, -1, -1, -1], [7, 5, 6, 4]]
    assert ec[0, 3] == float("inf") and ec[1, 1] == float("inf")
    # a padded result is safe to gather with: -1 gives a zero row
    rows = pool.gather_rows(torch.arange(16, dtype=torch.float32, device="cuda").reshape(8, 2), idx[1])
    assert rows.tolist() == [[6.0, 7.0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]]
    # bad arguments come back as errors, not launches
    with pytest.raises(RuntimeError, match="sc_select"):
        pool.select(c, 0)
    with pytest.raises((RuntimeError, ValueError)):
        pool.window(torch.zeros(3, 132, device="cuda"), torch.zeros(7, 68, device="cuda"), torch.zeros(132, device="cuda"), 1)


def test_two_handles_and_a_side_stream(character, pool, oracle):
    """Two pools (handles) on one device, one of them driven from a non-default stream: results equal the default
    path bit for bit (entry points enqueue on the caller's stream and share nothing between handles)."""
    import torch
    from samcon_b200.pool import GpuSamplePool
    rng = np.random.default_rng(12)
    parents = torch.tensor(np.array([stand(rng, 0.05, 0.2, 0.99) for _ in range(6)]), dtype=torch.float32, device="cuda")
    actions = torch.tensor(np.array([near_action(parents[k // 3].cpu().numpy(), rng, 0.1) for k in range(18)]),
                           dtype=torch.float32, device="cuda")
    target = torch.tensor(stand(rng), dtype=torch.float32, device="cuda")
    ref = pool.window(parents, actions, target, 30)
    other = GpuSamplePool(character, 0)
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            a = other.window(parents, actions, target, 30)
            idx, ec = other.select(a[1], 6)
        b = pool.window(parents, actions, target, 30)           # the first pool keeps working on the default stream
        side.synchronize()
        torch.cuda.synchronize()
        for x, y, z in zip(ref, a, b):
            assert torch.equal(x, y) and torch.equal(x, z)
        assert idx[0].tolist() == torch.argsort(ref[1], stable=True)[:6].tolist()
    finally:
        other.close()


def test_box_contacts_exact_narrow_phase(walk_cfg):
    """Link-vs-link pairs with a box on the GPU, SimParams(exact_box_contacts=True): exact box vs swept-segment narrow phase
    (of two boxes the second as its capsule): a foot box against the other shin, feet against each other."""
    from helpers import legs_crossed
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    from samcon_b200.pool import GpuSamplePool
    from test_emu_vs_oracle import _box_vs_round_poses
    ch = Character.from_config(walk_cfg, SimParams(exact_box_contacts=True))
    oracle = Oracle(ch)
    pool = GpuSamplePool(ch, 0)
    rng = np.random.default_rng(13)
    parents = np.concatenate([_box_vs_round_poses(oracle, 12), np.array([legs_crossed(rng, 0.10 + 0.01 * i, 2.0, 0.05) for i in range(4)])])
    actions = np.array([near_action(p, rng, 0.05) for p in parents])
    target = stand(rng)
    st, cost, info = _run(pool, parents, actions, target, 100, 1)
    ost, ocost, oinfo = oracle.window(parents, actions, target, 100)
    rms = joint_rms(st, ost)
    pool.close()
    print(f"box contact windows: joint rms median {np.median(rms):.2e} max {rms.max():.2e}")
    assert rms.max() < JOINT_TOL
    np.testing.assert_allclose(cost, ocost, rtol=COST_RTOL)


def test_first_principles_on_the_device(pool, character, oracle):
    """Two checks of the CUDA path that need no oracle for their expected value.  (i) Free flight: the whole-body linear
    momentum changes by gravity alone (first-order integrator: error O(dt)).  (ii) A humanoid standing still: the normal
    impulses the contact solve ends a sub-step with -- read from the contact cache -- carry the weight, sum = m g h."""
    import torch
    rng = np.random.default_rng(31)
    # (i)
    s0 = np.array([rand_state(rng, 1.0, 5.0) for _ in range(4)])
    acts = s0[:, 7:75].copy()
    d = torch.tensor(s0, dtype=torch.float32, device="cuda")
    f0 = pool.cost(d, d).cpu().numpy()                    # warm-up of the cost path; features are read through the oracle below
    pool.step(d, acts.astype(np.float32), 200)
    s1 = d.cpu().numpy().astype(np.float64)
    for a, b in zip(s0, s1):
        _, _, _, v0 = oracle.link_coms(a)
        _, _, _, v1 = oracle.link_coms(b)
        np.testing.assert_allclose(v1 - v0, [0.0, 0.0, -9.8 * 0.2], atol=0.06)      # PD torques are internal forces
    del f0
    # (ii)
    s = stand(rng, 0.0, 0.0, 0.99)
    s[0:2] = 0
    st = torch.tensor(s[None], dtype=torch.float32, device="cuda")
    cache = pool.empty_cache(1)
    act = s[None, 7:75].astype(np.float32)
    pool.step(st, act, 400, cache=cache)
    sums = []
    for _ in range(50):
        pool.step(st, act, 1, cache=cache)
        c = cache.cpu().numpy()[0]
        assert (c[:8] >= 0).sum() >= 4                    # both feet on the ground
        sums.append(c[8:].sum())
    h = character.sim.dt / character.sim.num_substeps
    weight = sum(l.mass for l in character.links) * 9.8
    assert abs(np.mean(sums) / h - weight) / weight < 0.01


def test_a_samples_result_does_not_depend_on_the_batch_around_it(pool):
    """The CTA's warps share the contact assembly (assembly_shared): who computes a column must not change a bit.  The same
    contact-rich samples are rolled out inside batches of different size and order -- different tiles, different numbers of
    tile and helper warps per CTA, one or several rounds -- and every state, cost and contact cache has to be identical."""
    from samcon_b200.config import Config
    from samcon_b200.sampling import get_batch_action
    rng = np.random.default_rng(11)
    cfg = Config("walk")
    n = 6000                                                       # 1500 tiles: 10-11 per CTA = two rounds
    base = np.array([stand(rng, 0.04, 0.3, 0.97 + 0.0005 * (i % 40)) for i in range(64)], np.float32)
    parents = base[rng.integers(0, 64, n)]
    target = stand(rng, 0.03, 0.1, 0.985)
    actions = get_batch_action(target, n, cfg.values, "uniform", np.random.RandomState(5)).astype(np.float32)
    ref = [x.cpu().numpy() for x in pool.window(parents, actions, target.astype(np.float32), 40, children=1, out_cache=True)]
    assert (ref[3][:, :8] >= 0).sum(1).max() >= 6                  # contact-rich: some samples hold 6+ contacts at the end
    for pick in (np.arange(37), np.arange(n)[::-1][:1203], rng.permutation(n)[:2500]):
        out = [x.cpu().numpy() for x in pool.window(parents[pick], actions[pick], target.astype(np.float32), 40, children=1, out_cache=True)]
        for a, b in zip(ref, out):
            assert np.array_equal(a[pick].view(np.uint32), b.view(np.uint32))


def test_assembly_pool_equals_the_private_assembly_bit_for_bit(pool, tmp_path):
    """The shared contact assembly, the helper warps and the reduced barrier set only move work between warps.  A build of
    the same sources with the pool compiled out (-DSC_SHARE_ASSEMBLY=0: every warp assembles its own tile and nobody helps)
    must produce the same bits -- states, costs, cost terms, contact caches -- on chained, contact-rich windows."""
    import subprocess
    from samcon_b200 import abi, build
    from samcon_b200.config import Config
    from samcon_b200.model import Character, SimParams
    from samcon_b200.pool import GpuSamplePool
    from samcon_b200.sampling import get_batch_action
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc on this box")
    lib = str(tmp_path / "libsamcon_b200_private.so")
    flags = [f for f in build.NVCC_FLAGS if f not in ("-Xptxas", "-v", "-lineinfo")]
    subprocess.check_call([nvcc, *flags, "-DSC_SHARE_ASSEMBLY=0", "-o", lib, os.path.join(os.path.dirname(abi.LIB_PATH), "samcon_b200.cu")])
    cfg = Config("walk")
    ch = Character.from_config(cfg, SimParams(self_collision=True))
    saved_path, saved_lib = abi.LIB_PATH, abi._lib
    try:
        abi.LIB_PATH, abi._lib = lib, None
        private = GpuSamplePool(ch, 0)
    finally:
        abi.LIB_PATH, abi._lib = saved_path, saved_lib
    try:
        rng = np.random.default_rng(21)
        S, E = 5000, 500                                           # 9 tile warps per CTA: the two-barrier instance of the loop
        parents = np.array([stand(rng, 0.04, 0.3, 0.975) for _ in range(E)], np.float32)
        caches = None
        target = stand(rng, 0.03, 0.1, 0.985).astype(np.float32)
        for w in range(3):
            actions = get_batch_action(target, S, cfg.values, "uniform", np.random.RandomState(7 + w)).astype(np.float32)
            a = pool.window(parents, actions, target, 50, parent_cache=caches, out_cache=True)
            b = private.window(parents, actions, target, 50, parent_cache=caches, out_cache=True)
            for x, y in zip(a, b):
                assert np.array_equal(x.cpu().numpy().view(np.uint32), y.cpu().numpy().view(np.uint32))
            idx, _ = pool.select(a[1], E)
            parents, caches = pool.gather_rows(a[0], idx.reshape(-1)), pool.gather_rows(a[3], idx.reshape(-1))
    finally:
        private.close()

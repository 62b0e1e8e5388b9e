"""Self-consistency of the CPU oracle (the reference ships no tests and no goldens for the physics -- SURVEY.md §4,
§8c -- so the oracle is pinned by the identities any correct restatement must satisfy)."""
import numpy as np
import pytest

from helpers import near_action, rand_state, stand


def _energy(o, ch, s):
    M = o.mass_matrix(s, 0)
    v = o.gen_vel(s)
    pos, _, _, _ = o.link_coms(s)
    m = np.array([l.mass for l in ch.links])
    return 0.5 * v @ M @ v + 9.8 * (m * pos[:, 2]).sum()


@pytest.mark.parametrize("iset", [0, 1])
def test_mass_matrix_spd_and_matches_rnea(oracle, iset):
    rng = np.random.default_rng(0)
    s = rand_state(rng)
    M = oracle.mass_matrix(s, iset)
    assert np.abs(M - M.T).max() < 1e-12 and np.linalg.eigvalsh(M).min() > 0
    assert abs(np.trace(M[3:6, 3:6]) / 3 - 53.5) < 1e-9           # total mass in the base linear block
    c0 = oracle.bias(s, iset, gravity=False)
    for i in range(0, 57, 4):                                     # M e_i = RNEA(q, v, e_i) - RNEA(q, v, 0)
        e = np.zeros(57)
        e[i] = 1
        assert np.abs(oracle.bias(s, iset, False, e) - c0 - M[:, i]).max() < 1e-10


def test_free_flight_conservation(oracle, character):
    """Zero torques, far above the ground: energy drift and momentum error are first-order in dt."""
    rng = np.random.default_rng(5)
    s = rand_state(rng, 1.0, 5.0)
    e0 = _energy(oracle, character, s)
    _, _, c0, v0 = oracle.link_coms(s)
    s2 = oracle.step_torque(s, np.zeros(57), 400)                 # 0.2 s of 0.5 ms sub-steps
    e1 = _energy(oracle, character, s2)
    _, _, c1, v1 = oracle.link_coms(s2)
    assert abs(e1 - e0) / abs(e0) < 5e-4
    g = np.array([0, 0, -9.8])
    assert np.abs(v1 - (v0 + 0.2 * g)).max() < 3e-3               # horizontal momentum kept, vertical follows g
    assert np.abs(c1 - (c0 + 0.2 * v0 + 0.5 * g * 0.04)).max() < 3e-3


def test_momentum_error_is_first_order_in_dt(walk_cfg):
    from samcon_b200.model import Character, SimParams
    from oracle.oracle import Oracle
    errs = []
    import copy
    for dt, n in ((1e-3, 400), (5e-4, 800)):
        cfg = copy.copy(walk_cfg)
        cfg.fps_sim = round(1.0 / dt)                                # the time step is the cfg's (setTimeStep(1/fps_sim))
        o = Oracle(Character.from_config(cfg))
        s = rand_state(np.random.default_rng(5), 1.0, 5.0)
        _, _, _, v0 = o.link_coms(s)
        _, _, _, v1 = o.link_coms(o.step_torque(s, np.zeros(57), n))
        errs.append(np.abs(v1 - v0 - 0.2 * np.array([0, 0, -9.8]))[:2].max())
    assert 1.7 < errs[0] / errs[1] < 2.3


def test_stable_pd_reduces_to_pd_and_clamps(oracle, character):
    rng = np.random.default_rng(1)
    s = rand_state(rng, 0.0, 3.0)
    s[75:] = 0
    a = near_action(s, rng, 0.02)
    tau = oracle.pd_torque(s, a)
    assert np.all(tau[:6] == 0)
    mf = np.repeat(character.max_forces, 3)
    assert np.all(np.abs(tau[6:]) <= mf + 1e-12)
    big = near_action(s, rng, 2.5)
    assert np.isclose(np.abs(oracle.pd_torque(s, big)[6:]) / mf, 1.0).any()      # saturates somewhere


def test_stable_pd_is_stable_at_reference_gains(oracle):
    """kp 500 / kd 50 at dt = 1 ms would blow up with explicit PD on the light links; stable PD must settle."""
    rng = np.random.default_rng(2)
    s = rand_state(rng, 0.0, 5.0, amp=0.2)
    s[75:] = 0
    a = near_action(s, rng, 0.1)
    for _ in range(3):
        s = oracle.step(s, a, 100)
    assert np.abs(s[81:]).max() < 1.0
    d = np.abs((s[7:75].reshape(17, 4) * a.reshape(17, 4)).sum(-1))
    assert (2 * np.arccos(np.clip(d, 0, 1))).max() < 0.1


def test_standing_contacts(oracle):
    rng = np.random.default_rng(3)
    s = stand(rng, 0.0, 0.0, 1.0)
    s[0:2] = 0
    a = s[7:75].copy()
    s = oracle.step(s, a, 200)
    c = oracle.contacts(s)
    assert len(c) == 8 and set(c[:, 0].astype(int)) == {3, 6}      # four corners of each foot box
    assert np.abs(c[:, 4]).max() < 1e-3                            # resting on the plane, no sinking
    assert 0.97 < s[2] < 1.0


def test_contact_cap_keeps_deepest(oracle, character):
    rng = np.random.default_rng(4)
    s = stand(rng, 0.0, 0.0, 0.95)                                  # feet 4 cm into the ground
    c = oracle.contacts(s)
    assert len(c) == character.sim.max_contacts
    assert (c[:, 4] < 0).all()


def test_cost_terms_against_numpy(oracle, character, walk_cfg):
    """/root/reference/envs/humanoid_tracking_env.py:78-176 written out with numpy on the oracle's link CoMs."""
    rng = np.random.default_rng(6)
    a, b = rand_state(rng, 1.0, 1.0), rand_state(rng, 1.0, 1.0)
    tot, (pose, root, ee, bal, com) = oracle.cost(a, b)
    pa, va, ca, cva = oracle.link_coms(a)
    pb, vb, cb, cvb = oracle.link_coms(b)

    def ang(q0, q1):
        return 2 * np.arccos(min(1.0, abs(np.dot(q0, q1))))
    jw = np.array(walk_cfg.joint_weights)
    p = sum(jw[j] * (ang(a[7 + 4 * j:11 + 4 * j], b[7 + 4 * j:11 + 4 * j]) ** 2
                     + 0.1 * np.sum((a[81 + 3 * j:84 + 3 * j] - b[81 + 3 * j:84 + 3 * j]) ** 2)) for j in range(17)) / 17
    r = ang(a[3:7], b[3:7]) ** 2 + 0.1 * np.sum((a[78:81] - b[78:81]) ** 2)
    ees = [e + 1 for e in walk_cfg.end_effectors]
    e_ = np.mean([(pa[i, 2] - pb[i, 2]) ** 2 for i in ees])
    bl = sum(np.sum((((ca - pa[i]) - (cb - pb[i]))[:2]) ** 2) for i in ees) / (4 * walk_cfg.height)
    cm = np.sum((ca - cb) ** 2) + 0.1 * np.sum((cva - cvb) ** 2)
    assert np.allclose([pose, root, ee, bal, com], [p, r, e_, bl, cm], rtol=1e-10)
    assert np.isclose(tot, np.dot(walk_cfg.cost_weights, [p, r, e_, bl, cm]))
    assert oracle.cost(a, a)[0] < 1e-12


def test_topk_ties_by_index(oracle):
    c = np.array([3.0, 1.0, 2.0, 1.0, 5.0, 1.0])
    assert list(oracle.topk(c, 4)) == [1, 3, 5, 2]


# ---------------------------------------------------------------------------------------------------------------
# link-vs-link contact (self_collision: True, walk.yml:113; flags humanoid_stable_pd.py:10-13)
# ---------------------------------------------------------------------------------------------------------------
def test_self_collision_pairs_exclude_ancestors(oracle, character):
    from samcon_b200.model import self_collision_pairs
    pairs = self_collision_pairs(character.links)
    assert oracle.num_pairs == len(pairs) == 125
    sh_link = [i for i, l in enumerate(character.links) for _ in l.shapes]
    names = [character.links[i].name for i in sh_link]
    pn = {frozenset((names[a], names[b])) for a, b in pairs}
    has = lambda a, b: frozenset((a, b)) in pn
    assert has("lknee", "rknee") and has("lankle", "rankle") and has("lelbow", "rhip") and has("lwrist", "rwrist")
    assert not any("root" in p for p in pn)                       # the root is everybody's ancestor
    assert not has("lhip", "lknee") and not has("chest", "lelbow") and not has("lelbow", "lwrist")


def test_self_contact_geometry(oracle, character):
    from helpers import legs_crossed
    rng = np.random.default_rng(0)
    assert len(oracle.contacts_full(legs_crossed(rng, 0.0))) == 0          # rest pose: nothing within 2 cm
    c = oracle.contacts_full(legs_crossed(rng, 0.12))
    assert len(c) >= 1 and (c[:, 1] >= 0).all()
    for r in c:
        n, pa, pb, d = r[8:11], r[2:5], r[5:8], r[11]
        assert abs(np.linalg.norm(n) - 1) < 1e-12
        assert np.allclose(pa - pb, d * n, atol=1e-12)             # the two surface points are d apart along the normal
        assert d < character.sim.contact_threshold
    names = {(character.links[int(r[0])].name, character.links[int(r[1])].name) for r in c}
    assert ("lknee", "rknee") in names


def test_self_contact_conserves_momentum_and_separates(walk_cfg):
    """No gravity, airborne, PD squeezing the legs together: contact impulses are internal, so linear momentum stays
    at the integrator's own (first-order, see test_momentum_error_is_first_order_in_dt) drift level; the legs must
    rest against each other instead of passing through to the PD target."""
    from helpers import legs_crossed
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    res = {}
    for sc in (True, False):
        ch = Character.from_config(walk_cfg, SimParams(gravity=(0.0, 0.0, 0.0), self_collision=sc))
        o = Oracle(ch)
        rng = np.random.default_rng(1)
        s = legs_crossed(rng, 0.05, 2.0)
        target = legs_crossed(rng, 0.35, 2.0)[7:75]
        m = np.array([l.mass for l in ch.links])

        def momentum(s):
            pos, vel, c, cv = o.link_coms(s)
            return m.sum() * cv

        p0 = momentum(s)
        deepest = 0.0
        for _ in range(30):
            s = o.step(s, target, 10)
            c = o.contacts_full(s)
            if len(c):
                deepest = min(deepest, c[:, 11].min())
        res[sc] = (np.abs(momentum(s) - p0).max(), deepest, s)
    assert res[True][0] < 5e-4 and res[True][0] < 2 * res[False][0] + 1e-5
    assert res[True][1] > -0.01            # with self collision the legs rest against each other
    hip = lambda s: 2 * np.arccos(abs(s[10]))
    assert hip(res[False][2]) > 0.3 and hip(res[True][2]) < 0.25   # without it they pass through to the PD target


def test_oracle_regression_fixtures():
    """tests/golden/oracle_windows.npz: the restatement's own outputs on seeded windows (a regression pin of the
    oracle, NOT a reference golden -- DESIGN.md §6).  Regenerate with tests/golden/make_oracle_golden.py."""
    import os
    from oracle.oracle import Oracle
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_windows.npz"))
    for key, cfgname in (("air", "walk"), ("stand", "walk"), ("self", "walk"), ("cartwheel", "cartwheel")):
        o = Oracle(Character.from_config(Config(cfgname)))
        st, cost, info = o.window(g[f"{key}_parents"], g[f"{key}_actions"], g[f"{key}_target"], int(g[f"{key}_nsteps"]))
        np.testing.assert_allclose(st, g[f"{key}_state"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(cost, g[f"{key}_cost"], rtol=1e-9)
        np.testing.assert_allclose(info, g[f"{key}_info"], rtol=1e-8, atol=1e-12)


def test_standing_ground_reaction_carries_the_weight(oracle, character):
    """First-principles check of dynamics + contact together: once a standing humanoid has settled, the ground's
    normal impulse per sub-step is weight x h (the CoM does not accelerate vertically) and the horizontal one stays a
    few per cent of it (the passive stand leans slowly), far inside the friction pyramid."""
    rng = np.random.default_rng(3)
    s = stand(rng, 0.0, 0.0, 0.99)
    s[0:2] = 0
    a = s[7:75].copy()
    cache = oracle.empty_cache()            # the contact cache travels with the state from call to call
    s = oracle.step(s, a, 400, cache)
    imp = []
    for _ in range(100):
        s = oracle.step(s, a, 1, cache)
        imp.append(oracle.last_ground_impulse())
    imp = np.array(imp)
    h = character.sim.dt / character.sim.num_substeps
    weight = sum(l.mass for l in character.links) * 9.8
    assert abs(imp[:, 2].mean() / h - weight) / weight < 0.01
    assert np.abs(imp[:, :2].mean(axis=0) / h).max() < 0.1 * weight < character.sim.friction * weight


def test_box_segment_closest_points_against_brute_force():
    """The exact box narrow phase (link-vs-link pairs with a foot / arm box): closest points of a box and a segment by
    safeguarded Newton on the convex piecewise-quadratic distance, against a dense scan; degenerate segments (spheres);
    segments through the box report the depth below the nearest face at the middle of the part inside."""
    import ctypes as C
    from oracle.oracle import lib
    L = lib()
    L.oc_box_segment.restype = C.c_double
    rng = np.random.default_rng(17)

    def call(h, p0, p1):
        q, p = np.zeros(3), np.zeros(3)
        ax = C.c_int(0)
        P = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.POINTER(C.c_double))
        d = L.oc_box_segment(P(h), P(p0), P(p1), q.ctypes.data_as(C.POINTER(C.c_double)), p.ctypes.data_as(C.POINTER(C.c_double)), C.byref(ax))
        return d, q, p, ax.value

    ts = np.linspace(0, 1, 20001)
    worst = 0.0
    for _ in range(300):
        h = rng.uniform(0.02, 0.15, 3)
        p0 = rng.uniform(-0.4, 0.4, 3)
        p1 = p0 if rng.random() < 0.2 else p0 + rng.uniform(-0.4, 0.4, 3)
        pts = p0[None] + ts[:, None] * (p1 - p0)[None]
        dist = np.linalg.norm(pts - np.clip(pts, -h, h), axis=1)
        d, q, p, ax = call(h, p0, p1)
        if dist.min() > 1e-9:
            assert ax == -1
            assert abs(d - dist.min()) < 2e-5 * max(1.0, dist.min()) + 1e-9                  # the scan's grid is the coarser one
            assert d <= dist.min() + 1e-8
            np.testing.assert_allclose(np.linalg.norm(p - q), d, atol=1e-12)
            assert np.all(np.abs(q) <= h + 1e-12)
            worst = max(worst, abs(d - dist.min()))
        else:
            inside = np.all(np.abs(pts) <= h, axis=1)
            assert ax >= 0 and d <= 0 and np.all(np.abs(p) <= h + 1e-9)
            np.testing.assert_allclose(-d, (h - np.abs(p)).min(), atol=1e-12)                # depth below the nearest face
            tm = 0.5 * (ts[inside][0] + ts[inside][-1])
            np.testing.assert_allclose(p, p0 + tm * (p1 - p0), atol=2e-4)                    # middle of the part inside
    assert worst < 1e-5


def test_exact_box_contacts_switch(walk_cfg):
    """exact_box_contacts (off by default): feet interpenetrating sideways (box-box: the first box exact, the second its capsule) and an arm
    box against the torso spheres give contacts whose points lie on the box surface; the capsule stand-in of round 1 is
    still there behind the switch."""
    from helpers import legs_crossed
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    rng = np.random.default_rng(9)
    s = legs_crossed(rng, 0.12, 2.0, 0.0)
    exact = Oracle(Character.from_config(walk_cfg, SimParams(exact_box_contacts=True))).contacts_full(s)
    standin = Oracle(Character.from_config(walk_cfg, SimParams(exact_box_contacts=False))).contacts_full(s)
    feet = lambda c: c[(c[:, 0] == 3) & (c[:, 1] == 6)]                       # lankle (link 3) vs rankle (link 6)
    assert len(feet(exact)) == 1 and len(feet(standin)) == 1
    assert abs(feet(exact)[0, 11] - feet(standin)[0, 11]) > 1e-3              # the two models differ by millimetres ...
    assert abs(feet(exact)[0, 11] - feet(standin)[0, 11]) < 2e-2              # ... not more
    other = lambda c: c[~((c[:, 0] == 3) & (c[:, 1] == 6))]
    np.testing.assert_allclose(other(exact), other(standin), atol=1e-12)      # capsule-capsule pairs are untouched

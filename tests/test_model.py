"""Model compiler: URDF parsing, built-in character vs the reference URDF, baked device tables."""
import os

import numpy as np
import pytest

from samcon_b200.data.humanoid_spec import emit_urdf
from samcon_b200.model import Character, link_inertia
from samcon_b200.urdf import J_FIXED, J_SPHERICAL, load_links, parse_urdf_text

REF_URDF = "/root/reference/data/character/humanoid.urdf"


def test_builtin_character_shape():
    links = parse_urdf_text(emit_urdf())
    assert len(links) == 20
    assert sum(l.jtype == J_SPHERICAL for l in links) == 17 and sum(l.jtype == J_FIXED for l in links) == 2
    assert abs(sum(l.mass for l in links) - 53.5) < 1e-12
    assert sum(len(l.shapes) for l in links) == 23


@pytest.mark.skipif(not os.path.isfile(REF_URDF), reason="reference tree not present")
def test_builtin_character_equals_reference_urdf():
    a = parse_urdf_text(emit_urdf())
    b = parse_urdf_text(open(REF_URDF).read())
    assert [l.name for l in a] == [l.name for l in b]
    for x, y in zip(a, b):
        assert x.mass == y.mass and x.parent == y.parent and x.jtype == y.jtype
        assert np.array_equal(x.com, y.com) and np.array_equal(x.jorg, y.jorg)
        assert len(x.shapes) == len(y.shapes)
        for s, t in zip(x.shapes, y.shapes):
            assert s.kind == t.kind and np.array_equal(s.dims, t.dims) and np.array_equal(s.pos, t.pos)
            assert np.allclose(s.rot, t.rot, atol=1e-15)


def test_model_file_resolution(tmp_path):
    assert len(load_links("data/character/humanoid.urdf")) == 20     # cfg value falls back to the built-in
    p = tmp_path / "x.urdf"
    p.write_text(emit_urdf())
    assert len(load_links(str(p))) == 20
    with pytest.raises(FileNotFoundError):
        load_links("nope/other.urdf")


def test_unsupported_urdf_features_raise():
    bad = emit_urdf().replace('type="spherical"', 'type="prismatic"', 1)
    with pytest.raises(NotImplementedError):
        parse_urdf_text(bad)


def test_device_tables(character):
    t = character.device_tables()
    assert t["nb"] == 18 and t["nlev"] == 7
    assert list(t["level_start"]) == [0, 1, 4, 7, 10, 13, 16, 18]
    assert (np.diff(t["level_start"]) <= 4).all()
    # parents precede children, bodies of a level are consecutive
    assert all(t["parent"][b] < b for b in range(1, 18))
    assert sorted(t["jidx"][1:]) == list(range(17))
    # folded wrists: elbow bodies carry elbow + wrist mass
    names = [character.links[i].name for i in t["body_link"]]
    for side in "lr":
        b = names.index(side + "elbow")
        assert abs(t["inertia_dyn"][b, 0] - 1.5) < 1e-12 and abs(t["inertia_pd"][b, 0] - 1.5) < 1e-12
    assert abs(t["inertia_dyn"][:, 0].sum() - 53.5) < 1e-9
    assert len(t["cand_body"]) == 9 + 2 * 8 + 8 * 6           # spheres, capsule caps, box corners
    assert len(t["cp_body"]) == 20 and list(t["ee"]) == [15, 19, 3, 6]


def test_inertia_models(character):
    lank = next(l for l in character.links if l.name == "lankle")
    c1, I1 = link_inertia(lank, "aabb")
    c2, I2 = link_inertia(lank, "first_shape")
    assert np.allclose(c1, c2) and np.allclose(I1, I2)           # axis-aligned box: both models coincide
    x, y, z = 0.0875, 0.06, 0.185
    assert np.allclose(np.diag(I1), 1.0 / 12 * np.array([y * y + z * z, x * x + z * z, x * x + y * y]))
    root = character.links[0]
    c, I = link_inertia(root, "first_shape")
    assert np.allclose(c, [0.00354, 0.065, -0.03107]) and np.allclose(np.diag(I), 0.4 * 5 * 0.05 ** 2)


def test_character_validation(walk_cfg):
    with pytest.raises(ValueError):
        Character(load_links(None), np.ones(16), np.ones(17), np.ones(17), np.ones(17), [14, 18, 2, 5], np.ones(5), 1.65)


def test_time_step_follows_cfg_fps_sim(walk_cfg):
    """setTimeStep(1 / cfg.fps_sim) (humanoid_tracking_env.py:25, 46): the character built from a cfg simulates at the
    cfg's rate (it used to keep the 1 ms default), and a contradicting explicit SimParams is refused."""
    import copy
    import numpy as np
    import pytest
    from helpers import near_action, stand
    from oracle.oracle import Oracle
    from samcon_b200.model import Character, SimParams
    cfg = copy.copy(walk_cfg)
    cfg.fps_sim = 500
    ch = Character.from_config(cfg)
    assert abs(ch.sim.dt - 1.0 / 500) < 1e-15
    with pytest.raises(ValueError, match="fps_sim"):
        Character.from_config(cfg, SimParams())                   # dt = 1/1000 against fps_sim = 500
    assert Character.from_config(cfg, SimParams(dt=1.0 / 500)).sim.dt == 1.0 / 500
    # 50 steps at 500 Hz and 100 steps at 1000 Hz cover the same 0.1 s: a free fall ends at the same height
    rng = np.random.default_rng(0)
    s = stand(rng, 0.0, 0.0, 3.0)
    a = near_action(s, rng, 0.0)
    z500 = Oracle(ch).step(s, a, 50)[2]
    z1000 = Oracle(Character.from_config(walk_cfg)).step(s, a, 100)[2]
    assert abs((3.0 - z500) - 0.5 * 9.8 * 0.01) < 2e-3 and abs(z500 - z1000) < 5e-4      # semi-implicit Euler: O(dt) apart

"""An independent check of the model compiler's tables: kernel and oracle both take
their masses, centres of mass and inertias from samcon_b200/model.py, so an error there is common-mode and invisible to
every kernel-vs-oracle test.  Here the same quantities are derived a second time, in this file, straight from the URDF
text (the reference's data/character/humanoid.urdf when the checkout is present, else the bundled copy of the same
numbers) with hand-written formulas, and compared with what model.py hands to the oracle (raw_tables) and to the device
(device_tables)."""
import math
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

REF_URDF = "/root/reference/data/character/humanoid.urdf"


def _urdf_text():
    if os.path.isfile(REF_URDF):
        return open(REF_URDF).read()
    from samcon_b200.data.humanoid_spec import emit_urdf
    return emit_urdf()


def _parse(text):
    """name -> dict(mass, com, parent, jorg, shapes[(kind, dims, xyz, rpy)]) with nothing but ElementTree."""
    root = ET.fromstring(text)
    v = lambda s: np.array([float(x) for x in s.split()])
    links = {}
    order = []
    for l in root.findall("link"):
        ine = l.find("inertial")
        shapes = []
        for c in l.findall("collision"):
            o = c.find("origin")
            g = list(c.find("geometry"))[0]
            dims = {"sphere": lambda: (float(g.get("radius")),), "capsule": lambda: (float(g.get("radius")), float(g.get("length"))),
                    "box": lambda: tuple(v(g.get("size")))}[g.tag]()
            shapes.append((g.tag, dims, v(o.get("xyz")), v(o.get("rpy"))))
        links[l.get("name")] = {"mass": float(ine.find("mass").get("value")), "com": v(ine.find("origin").get("xyz")),
                                "parent": None, "jorg": np.zeros(3), "shapes": shapes}
        order.append(l.get("name"))
    for j in root.findall("joint"):
        c = links[j.find("child").get("link")]
        c["parent"] = j.find("parent").get("link")
        c["jorg"] = v(j.find("origin").get("xyz"))
    return links, order


def _rpy(r, p, y):                    # URDF fixed-axis roll-pitch-yaw: R = Rz(y) Ry(p) Rx(r)
    Rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    Ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    Rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _about_origin(m, c, Ic):          # parallel axis: inertia about the link origin from the one about the centre c
    return Ic + m * (np.dot(c, c) * np.eye(3) - np.outer(c, c))


def test_masses_and_zero_pose_com(character):
    links, order = _parse(_urdf_text())
    assert len(order) == 20 and abs(sum(l["mass"] for l in links.values()) - 53.5) < 1e-12
    # zero pose: every rotation is the identity, link origins are sums of joint origins
    pos = {}
    for n in order:
        l = links[n]
        pos[n] = (pos[l["parent"]] if l["parent"] else np.zeros(3)) + l["jorg"]
    com = sum(l["mass"] * (pos[n] + l["com"]) for n, l in links.items()) / 53.5
    # the same number through the oracle's tables and through the device tables
    from oracle.oracle import Oracle
    s = np.zeros(132); s[6] = 1.0; s[10:75:4] = 1.0
    _, _, ocom, _ = Oracle(character).link_coms(s)
    np.testing.assert_allclose(ocom, com, atol=1e-12)
    t = character.device_tables()
    assert abs(t["inertia_dyn"][:, 0].sum() - 53.5) < 1e-9 and abs(t["inertia_pd"][:, 0].sum() - 53.5) < 1e-9
    body_pos = np.zeros((t["nb"], 3))
    for b in range(1, t["nb"]):
        body_pos[b] = body_pos[t["parent"][b]] + t["jorg"][b]
    dcom = (t["inertia_dyn"][:, 0:1] * body_pos + t["inertia_dyn"][:, 1:4]).sum(0) / 53.5         # h = m c per body
    np.testing.assert_allclose(dcom, com, atol=1e-12)
    # both wrists are folded into their elbows on the device: 18 moving bodies carry the 20 links' mass
    assert t["nb"] == 18


def test_two_links_inertia_by_hand(character):
    """lankle (a box) and lknee (a rotated capsule), both inertia models, against model.py's oracle and device tables."""
    links, order = _parse(_urdf_text())
    raw = character.raw_tables()
    dev = character.device_tables()
    body_of = {int(l): b for b, l in enumerate(dev["body_link"])}

    def dev_Io(key, b):
        r = dev[key][b]
        xx, xy, xz, yy, yz, zz = r[4:10]
        return r[0], r[1:4], np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])

    # ---- lankle: box 0.0875 x 0.06 x 0.185, 1 kg, unrotated, centred on the link's CoM: both models give the solid box
    i = order.index("lankle")
    (kind, (sx, sy, sz), xyz, rpy), = links["lankle"]["shapes"]
    assert kind == "box" and np.allclose(rpy, 0) and np.allclose(xyz, links["lankle"]["com"])
    Ibox = np.diag([sy * sy + sz * sz, sx * sx + sz * sz, sx * sx + sy * sy]) * 1.0 / 12.0
    np.testing.assert_allclose(np.diag(Ibox), [0.0031520833, 0.0034901042, 0.0009380208], rtol=1e-6)    # typed in by hand
    for key in ("dyn", "pd"):
        np.testing.assert_allclose(raw["I_" + key][i].reshape(3, 3), Ibox, atol=1e-12)
        np.testing.assert_allclose(raw["com_" + key][i], links["lankle"]["com"], atol=1e-12)
        m, h, Io = dev_Io("inertia_" + key, body_of[i])
        assert m == 1.0
        np.testing.assert_allclose(h, links["lankle"]["com"], atol=1e-12)
        np.testing.assert_allclose(Io, _about_origin(1.0, links["lankle"]["com"], Ibox), atol=1e-12)

    # ---- lknee: capsule r = 0.05, L = 0.325, 3 kg, rotated by rpy
    i = order.index("lknee")
    (kind, (r, L), xyz, rpy), = links["lknee"]["shapes"]
    assert kind == "capsule" and (r, L) == (0.05, 0.325)
    R = _rpy(*rpy)
    m = 3.0
    # (a) stable-PD model: uniform solid capsule = cylinder + two hemispheres, axis = shape z
    vc, vs = math.pi * r * r * L, 4.0 / 3.0 * math.pi * r ** 3
    mc = m * vc / (vc + vs); ms = m - mc
    i_ax = 0.5 * mc * r * r + 0.4 * ms * r * r
    i_pe = mc * (L * L / 12.0 + r * r / 4.0) + ms * (0.4 * r * r + 0.25 * L * L + 0.375 * L * r)
    np.testing.assert_allclose([i_ax, i_pe], [0.0036223, 0.040574], rtol=2e-5)                          # worked out by hand
    Ipd = R @ np.diag([i_pe, i_pe, i_ax]) @ R.T
    np.testing.assert_allclose(raw["I_pd"][i].reshape(3, 3), Ipd, atol=1e-12)
    mm, h, Io = dev_Io("inertia_pd", body_of[i])
    np.testing.assert_allclose(Io, _about_origin(m, xyz, Ipd), atol=1e-12)
    np.testing.assert_allclose(h, m * xyz, atol=1e-12)
    # (b) dynamics model: box inertia of the capsule's axis-aligned bounding box in the link frame
    half = np.abs(R) @ np.array([r, r, r + 0.5 * L])
    lx, ly, lz = 2 * half
    Idyn = m / 12.0 * np.diag([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
    np.testing.assert_allclose(raw["I_dyn"][i].reshape(3, 3), Idyn, atol=1e-12)
    mm, h, Io = dev_Io("inertia_dyn", body_of[i])
    np.testing.assert_allclose(Io, _about_origin(m, links["lknee"]["com"], Idyn), atol=1e-12)
    # the capsule stands nearly along the link's -y (links hang along -y): its bounding box is tallest there
    assert ly > 0.4 and lx < 0.13 and lz < 0.16


@pytest.mark.skipif(not os.path.isfile(REF_URDF), reason="reference checkout absent")
def test_bundled_character_equals_reference_urdf():
    from samcon_b200.data.humanoid_spec import emit_urdf
    a, oa = _parse(open(REF_URDF).read())
    b, ob = _parse(emit_urdf())
    assert oa == ob
    for n in oa:
        assert a[n]["mass"] == b[n]["mass"] and a[n]["parent"] == b[n]["parent"]
        np.testing.assert_array_equal(a[n]["com"], b[n]["com"])
        np.testing.assert_array_equal(a[n]["jorg"], b[n]["jorg"])
        assert len(a[n]["shapes"]) == len(b[n]["shapes"])
        for (ka, da, xa, ra), (kb, db, xb, rb) in zip(a[n]["shapes"], b[n]["shapes"]):
            assert ka == kb and tuple(da) == tuple(db)
            np.testing.assert_array_equal(xa, xb); np.testing.assert_array_equal(ra, rb)

"""Minimal URDF reader for SamCon characters.

Replaces what the reference obtains from ``pybullet.loadURDF``
(/root/reference/envs/humanoid_stable_pd.py:9-20): link masses, inertial
origins, collision primitives and the joint tree, in URDF order
(URDF_MAINTAIN_LINK_ORDER).  Only what the hot path needs is supported:
floating root, ``spherical`` and ``fixed`` joints with zero joint rpy, and
sphere / capsule / box collision primitives.  Anything else raises, like the
reference does for unknown joint types (humanoid_stable_pd.py:40-41).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

SPHERE, CAPSULE, BOX = 0, 1, 2
J_FLOATING, J_SPHERICAL, J_FIXED = 0, 1, 2


def rpy_to_matrix(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix (R = Rz Ry Rx)."""
    r, p, y = (float(a) for a in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]], dtype=np.float64)


@dataclass
class Shape:
    kind: int
    dims: np.ndarray          # (3,) sphere: r,0,0  capsule: r,L,0  box: full extents
    pos: np.ndarray           # (3,) in link frame
    rot: np.ndarray           # (3,3) shape axes in link frame


@dataclass
class Link:
    name: str
    mass: float
    com: np.ndarray           # inertial origin in link frame
    parent: int               # -1 for the root
    jtype: int
    jorg: np.ndarray          # joint origin in the parent link frame
    shapes: list = field(default_factory=list)


def _vec(s, n=3):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if v.shape != (n,):
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return v


def parse_urdf_text(text: str, scale: float = 1.0) -> list[Link]:
    root = ET.fromstring(text)
    raw = {}
    order = []
    for l in root.findall("link"):
        name = l.get("name")
        ine = l.find("inertial")
        if ine is None:
            raise ValueError(f"link {name}: no <inertial>")
        org = ine.find("origin")
        com = _vec(org.get("xyz", "0 0 0")) if org is not None else np.zeros(3)
        if org is not None and np.any(_vec(org.get("rpy", "0 0 0")) != 0):
            raise NotImplementedError(f"link {name}: rotated inertial frame")
        mass = float(ine.find("mass").get("value"))
        shapes = []
        for c in l.findall("collision"):
            corg = c.find("origin")
            pos = _vec(corg.get("xyz", "0 0 0")) if corg is not None else np.zeros(3)
            rot = rpy_to_matrix(_vec(corg.get("rpy", "0 0 0"))) if corg is not None else np.eye(3)
            g = list(c.find("geometry"))[0]
            if g.tag == "sphere":
                kind, dims = SPHERE, np.array([float(g.get("radius")), 0.0, 0.0])
            elif g.tag == "capsule":
                kind, dims = CAPSULE, np.array([float(g.get("radius")), float(g.get("length")), 0.0])
            elif g.tag == "box":
                kind, dims = BOX, _vec(g.get("size"))
            else:
                raise NotImplementedError(f"link {name}: collision geometry <{g.tag}>")
            # globalScaling (humanoid_stable_pd.py:17) scales every length
            shapes.append(Shape(kind, dims * scale, pos * scale, rot))
        # PyBullet scales mass by scale^3 under globalScaling [BULLET-UPSTREAM]; scale==1 in both cfgs
        raw[name] = Link(name, mass * scale ** 3, com * scale, -1, J_FLOATING, np.zeros(3), shapes)
        order.append(name)
    for j in root.findall("joint"):
        child = j.find("child").get("link")
        parent = j.find("parent").get("link")
        jt = j.get("type")
        if jt == "spherical":
            t = J_SPHERICAL
        elif jt == "fixed":
            t = J_FIXED
        else:
            raise NotImplementedError(f"joint {j.get('name')}: type {jt}")
        org = j.find("origin")
        if org is not None and np.any(_vec(org.get("rpy", "0 0 0")) != 0):
            raise NotImplementedError(f"joint {j.get('name')}: rotated joint frame")
        raw[child].jtype = t
        raw[child].jorg = (_vec(org.get("xyz", "0 0 0")) if org is not None else np.zeros(3)) * scale
        raw[child].parent = order.index(parent)
    links = [raw[n] for n in order]
    roots = [i for i, l in enumerate(links) if l.parent < 0]
    if roots != [0]:
        raise ValueError("the first URDF link must be the single root")
    for i, l in enumerate(links):
        if i and l.parent >= i:
            raise ValueError("links must be listed parents-first")
    return links


def load_links(model_file: str | None, scale: float = 1.0, search_dirs=()) -> list[Link]:
    """Resolve the cfg key ``model_file`` (walk.yml) to a list of links.

    A readable path wins; otherwise a file named ``humanoid.urdf`` falls back
    to the built-in character table (same numbers as the reference URDF).
    """
    cands = []
    if model_file:
        cands.append(model_file)
        cands += [os.path.join(d, model_file) for d in search_dirs]
    for c in cands:
        if os.path.isfile(c):
            with open(c, "r") as f:
                return parse_urdf_text(f.read(), scale)
    if model_file is None or os.path.basename(model_file) == "humanoid.urdf":
        from .data.humanoid_spec import emit_urdf
        return parse_urdf_text(emit_urdf(), scale)
    raise FileNotFoundError(model_file)

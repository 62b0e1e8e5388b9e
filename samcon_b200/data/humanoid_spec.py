"""Numeric specification of the SMPL-like humanoid used by SamCon.

The reference ships this character as ``data/character/humanoid.urdf``
(/root/reference/data/character/humanoid.urdf:5-491; table in SURVEY.md §9.1).
Here the same physical parameters are held as a compact table and turned
into URDF text on demand (``emit_urdf``), so that the single URDF parser in
``samcon_b200.urdf`` is the only code path a model ever goes through -- a
user-supplied URDF and the built-in character are treated identically.

Units: metres, kilograms, radians.  Link order == URDF order, which is the
order PyBullet keeps under URDF_MAINTAIN_LINK_ORDER
(/root/reference/envs/humanoid_stable_pd.py:9), i.e. link index = row - 1.
"""

# shape tuples: (kind, dims, xyz, rpy)
#   sphere : dims = (radius,)
#   capsule: dims = (radius, length)      axis = local Z of the shape frame
#   box    : dims = (sx, sy, sz)          full extents
S, C, B = "sphere", "capsule", "box"

# name, mass, com xyz, parent, joint type, joint origin xyz, shapes
LINKS = [
    ("root", 5.0, (0, 0, 0), None, "floating", (0, 0, 0), [
        (S, (0.05,), (0.00354, 0.065, -0.03107), (0, 1.5708, 0)),
        (S, (0.075,), (-0.05769, -0.02577, -0.0174), (0, 0, 0)),
        (S, (0.075,), (0.06735, -0.02415, -0.0174), (0, 0, 0))]),
    ("lhip", 5.0, (0.02173, -0.19323, 0.00402), "root", "spherical", (0.08858, -0.08228, -0.01766), [
        (C, (0.05, 0.3), (0.02173, -0.19323, 0.00402), (1.55, 0.11194, 0.10964))]),
    ("lknee", 3.0, (-0.00739, -0.21344, -0.01871), "lhip", "spherical", (0.04345, -0.38647, 0.00804), [
        (C, (0.05, 0.325), (-0.00739, -0.21344, -0.01871), (1.65825, -0.0345, -0.03766))]),
    ("lankle", 1.0, (0.01719, -0.06032, 0.02617), "lknee", "spherical", (-0.01479, -0.42687, -0.03743), [
        (B, (0.0875, 0.06, 0.185), (0.01719, -0.06032, 0.02617), (0, 0, 0))]),
    ("rhip", 5.0, (-0.02163, -0.19184, -0.00242), "root", "spherical", (-0.09031, -0.09051, -0.01354), [
        (C, (0.05, 0.3), (-0.02163, -0.19184, -0.00242), (1.58342, -0.11226, -0.11368))]),
    ("rknee", 3.0, (0.00953, -0.21002, -0.01728), "rhip", "spherical", (-0.04326, -0.38369, -0.00484), [
        (C, (0.05, 0.325), (0.00953, -0.21002, -0.01728), (1.65289, 0.04518, 0.04905))]),
    ("rankle", 1.0, (-0.01742, -0.06202, 0.02969), "rknee", "spherical", (0.01906, -0.42005, -0.03456), [
        (B, (0.0875, 0.06, 0.185), (-0.01742, -0.06202, 0.02969), (0, 0, 0))]),
    ("lowerback", 5.0, (0.0, 0.05, 0.013), "root", "spherical", (0.0, 0.1244, -0.03), [
        (S, (0.065,), (0.0, 0.06, 0.013), (0, 1.5708, 0))]),
    ("upperback", 5.0, (0.0, 0.02246, 0.00143), "lowerback", "spherical", (0.0, 0.13796, 0.02682), [
        (S, (0.05,), (0.0, 0.03246, 0.00143), (0, 1.5708, 0))]),
    ("chest", 8.0, (0.0, 0.057, -0.00687), "upperback", "spherical", (0.0, 0.05603, 0.00285), [
        (S, (0.07,), (-0.045, 0.057, -0.00687), (0, 0, 0)),
        (S, (0.07,), (0.045, 0.057, -0.00687), (0, 0, 0))]),
    ("lowerneck", 0.5, (0.0, -0.01296, 0.01), "chest", "spherical", (0.0, 0.15524, -0.03347), [
        (C, (0.030, 0.0273), (0.0, -0.01296, 0.01), (1.5708, 0, 0))]),
    ("upperneck", 3.0, (0, 0.0020, 0), "lowerneck", "spherical", (0.0, 0.08894, 0.02041), [
        (C, (0.06, 0.035), (0, 0.0020, 0), (1.5708, 0, 0))]),
    ("lclavicle", 1.0, (0.06146, 0.0226, -0.00952), "chest", "spherical", (0.0717, 0.114, -0.0189), [
        (C, (0.04, 0.05), (0.06146, 0.0226, -0.00952), (1.17204, 1.95049, 1.5502))]),
    ("lshoulder", 2.0, (0.12767, 0.0, 0.0), "lclavicle", "spherical", (0.12292, 0.04521, -0.01905), [
        (B, (0.25, 0.05, 0.05), (0.12767, 0.0, 0.0), (0, 0, 0))]),
    ("lelbow", 1.0, (0.12285, 0.0, 0.0), "lshoulder", "spherical", (0.25533, 0.0, 0.0), [
        (B, (0.22, 0.05, 0.05), (0.12285, 0.0, 0.0), (0, 0, 0))]),
    ("lwrist", 0.5, (0.02228, 0, 0), "lelbow", "fixed", (0.24571, 0.0, 0.0), [
        (S, (0.04,), (0.02228, 0, 0), (0, 0, 0))]),
    ("rclavicle", 1.0, (-0.05661, 0.02343, -0.00424), "chest", "spherical", (-0.08295, 0.11247, -0.02371), [
        (C, (0.04, 0.05), (-0.05661, 0.02343, -0.00424), (-1.74968, -1.17274, 1.34362))]),
    ("rshoulder", 2.0, (-0.13006, 0.0, 0.0), "rclavicle", "spherical", (-0.11323, 0.04685, -0.00847), [
        (B, (0.25, 0.05, 0.05), (-0.13006, 0.0, 0.0), (0, 0, 0))]),
    ("relbow", 1.0, (-0.12455, 0.0, 0.0), "rshoulder", "spherical", (-0.26013, 0.0, 0.0), [
        (B, (0.22, 0.05, 0.05), (-0.12455, 0.0, 0.0), (0, 0, 0))]),
    ("rwrist", 0.5, (-0.0239, 0, 0), "relbow", "fixed", (-0.24911, 0.0, 0.0), [
        (S, (0.04,), (-0.0239, 0, 0), (0, 0, 0))]),
]


def _f(v):
    return " ".join(repr(float(x)) for x in v)


def emit_urdf(name="smpl_humanoid"):
    """Return URDF text for the built-in character."""
    out = ['<?xml version="1.0"?>', f'<robot name="{name}">']
    for lname, mass, com, _p, _jt, _jo, shapes in LINKS:
        out.append(f'<link name="{lname}"><inertial><origin xyz="{_f(com)}" rpy="0 0 0"/>'
                   f'<mass value="{mass!r}"/>'
                   '<inertia ixx="0.001" ixy="0" ixz="0" iyy="0.001" iyz="0" izz="0.001"/></inertial>')
        for k, (kind, dims, xyz, rpy) in enumerate(shapes):
            if kind == S:
                geo = f'<sphere radius="{dims[0]!r}"/>'
            elif kind == C:
                geo = f'<capsule radius="{dims[0]!r}" length="{dims[1]!r}"/>'
            else:
                geo = f'<box size="{_f(dims)}"/>'
            out.append(f'<collision name="c{k}_{lname}"><origin xyz="{_f(xyz)}" rpy="{_f(rpy)}"/>'
                       f'<geometry>{geo}</geometry></collision>')
        out.append('</link>')
    for lname, _m, _c, parent, jtype, jorg, _s in LINKS:
        if parent is None:
            continue
        out.append(f'<joint name="{lname}" type="{jtype}"><parent link="{parent}"/><child link="{lname}"/>'
                   f'<origin xyz="{_f(jorg)}" rpy="0 0 0"/></joint>')
    out.append('</robot>')
    return "\n".join(out) + "\n"

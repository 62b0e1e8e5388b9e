"""ctypes mirror of include/samcon_b200.h and loader of the in-tree CUDA library.

There is no CPU fallback: if ``libsamcon_b200.so`` is missing or does not export every symbol the header
declares, importing the pool raises.  (The library is built by ``__graft_entry__.build()`` /
``python -m samcon_b200.build`` with nvcc for sm_100a and lives next to the sources so that it travels
with the repo snapshot.)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB_PATH = os.path.join(_CSRC, "libsamcon_b200.so")

SC_STATE_DIM = 132
SC_MAX_BODIES = 18
SC_MAX_CONTACTS = 8
SC_CACHE_DIM = 2 * SC_MAX_CONTACTS

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class ScModel(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("nlev", C.c_int32), ("ncand", C.c_int32), ("ncp", C.c_int32), ("nee", C.c_int32),
        ("parent", i32p), ("jidx", i32p), ("level_start", i32p), ("cand_body", i32p), ("cp_body", i32p), ("ee", i32p),
        ("jorg", f32p), ("inertia_dyn", f32p), ("inertia_pd", f32p), ("kp", f32p), ("kd", f32p), ("maxf", f32p),
        ("cand", f32p), ("cp_off", f32p), ("cp_mass", f32p), ("joint_weights", f32p),
        ("cost_weights", C.c_float * 5), ("height", C.c_float), ("dt", C.c_float),
        ("num_substeps", C.c_int32), ("num_solver_iters", C.c_int32), ("max_contacts", C.c_int32),
        ("gravity", C.c_float * 3), ("friction", C.c_float), ("erp", C.c_float), ("contact_threshold", C.c_float),
        ("max_velocity", C.c_float),
        ("nprim", C.c_int32), ("npair", C.c_int32), ("prim_body", i32p), ("prim", f32p), ("pair", i32p),
        ("friction_self", C.c_float), ("warmstart", C.c_float), ("spinning_friction", C.c_float),
        ("spinning_friction_self", C.c_float),
        ("nbox", C.c_int32), ("prim_box", i32p), ("box", f32p),
    ]


def make_model(character):
    """Character -> (ScModel, keepalive list).  Arrays are float32/int32 host copies."""
    t = character.device_tables()
    sp = character.sim
    keep = []

    def F(a):
        a = np.ascontiguousarray(a, dtype=np.float32); keep.append(a)
        return a.ctypes.data_as(f32p)

    def I(a):
        a = np.ascontiguousarray(a, dtype=np.int32); keep.append(a)
        return a.ctypes.data_as(i32p)
    nb = t["nb"]
    jw = np.zeros(nb)
    for b in range(1, nb):
        jw[b] = t["joint_weights"][t["jidx"][b]]
    m = ScModel()
    m.nb, m.nlev, m.ncand, m.ncp, m.nee = nb, t["nlev"], len(t["cand_body"]), len(t["cp_body"]), len(t["ee"])
    m.parent, m.jidx, m.level_start = I(t["parent"]), I(t["jidx"]), I(t["level_start"])
    m.cand_body, m.cp_body, m.ee = I(t["cand_body"]), I(t["cp_body"]), I(t["ee"])
    m.jorg, m.inertia_dyn, m.inertia_pd = F(t["jorg"]), F(t["inertia_dyn"]), F(t["inertia_pd"])
    m.kp, m.kd, m.maxf = F(t["kp"]), F(t["kd"]), F(t["maxf"])
    m.cand = F(np.concatenate([t["cand_pt"], t["cand_r"][:, None]], axis=1))
    m.cp_off, m.cp_mass, m.joint_weights = F(t["cp_off"]), F(t["cp_mass"]), F(jw)
    m.cost_weights = (C.c_float * 5)(*[float(x) for x in t["cost_weights"]])
    m.height, m.dt = float(t["height"]), float(sp.dt)
    m.num_substeps, m.num_solver_iters, m.max_contacts = sp.num_substeps, sp.num_solver_iters, sp.max_contacts
    m.gravity = (C.c_float * 3)(*[float(x) for x in sp.gravity])
    m.friction, m.erp, m.contact_threshold, m.max_velocity = sp.friction, sp.erp, sp.contact_threshold, sp.max_velocity
    m.nprim = len(t["prim_body"])
    m.npair = len(t["pair"]) if sp.self_collision else 0
    m.prim_body, m.prim = I(t["prim_body"]), F(t["prim"])
    m.pair = I(t["pair"] if len(t["pair"]) else np.zeros((1, 2)))
    m.friction_self = sp.friction_self
    m.warmstart, m.spinning_friction, m.spinning_friction_self = sp.warmstart, sp.spinning_friction, sp.spinning_friction_self
    m.nbox = len(t["box"])
    m.prim_box = I(t["prim_box"])
    m.box = F(t["box"] if len(t["box"]) else np.zeros((1, 15)))
    return m, keep


# every entry point include/samcon_b200.h declares
SYMBOLS = ["sc_create", "sc_destroy", "sc_last_error", "sc_sample_gen", "sc_window", "sc_step", "sc_cost",
           "sc_select", "sc_gather_rows", "sc_launch_count", "sc_rollout_config", "sc_contact_stats", "sc_motion_states", "sc_gather_rows_peer",
           "sc_measure_fp32_peak"]

_lib = None


def load():
    """Load libsamcon_b200.so and declare prototypes; raises if it is absent or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built -- run `python -m samcon_b200.build` (nvcc, sm_100a). "
                           "There is no CPU fallback for the SamCon rollout path.")
    lib = C.CDLL(LIB_PATH)
    for s in SYMBOLS:
        if not hasattr(lib, s):
            raise RuntimeError(f"{LIB_PATH} does not export {s}")
    vp, i, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.sc_create.argtypes = [C.POINTER(vp), C.POINTER(ScModel), i]
    lib.sc_destroy.argtypes = [vp]
    lib.sc_last_error.argtypes = [vp]
    lib.sc_last_error.restype = C.c_char_p
    lib.sc_sample_gen.argtypes = [vp, vp, i, vp, vp, vp, i, u64, u64, i, vp, vp]
    lib.sc_motion_states.argtypes = [vp, vp, vp, vp, i, vp, i, vp, vp]
    lib.sc_window.argtypes = [vp, vp, vp, i, vp, i, vp, vp, i, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.sc_step.argtypes = [vp, vp, vp, vp, i, i, vp]
    lib.sc_contact_stats.argtypes = [vp, C.POINTER(C.c_int64), i]
    lib.sc_cost.argtypes = [vp, vp, vp, i, vp, vp]
    lib.sc_select.argtypes = [vp, vp, vp, i, i, vp, vp, i, vp]
    lib.sc_gather_rows.argtypes = [vp, vp, vp, i, i, vp, vp]
    lib.sc_gather_rows_peer.argtypes = [vp, vp, vp, vp, i, i, vp, vp]
    lib.sc_measure_fp32_peak.argtypes = [i, C.POINTER(C.c_double)]
    lib.sc_launch_count.argtypes = [vp]
    lib.sc_launch_count.restype = C.c_int64
    lib.sc_rollout_config.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    _lib = lib
    return lib

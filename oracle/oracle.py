"""ctypes front-end of oracle/samcon_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
PARITY UNPINNED (no PyBullet, no reference goldens for the physics) -- see the C file header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libsamcon_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "samcon_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsamcon_oracle.so"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oc_create.restype = C.c_void_p
        _lib.oc_nv.argtypes = [C.c_void_p]
        _lib.oc_destroy.argtypes = [C.c_void_p]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


class Oracle:
    """One humanoid + engine settings; methods mirror the reference env's calls."""

    def __init__(self, character):
        L = lib()
        t = character.raw_tables()
        sp = character.sim
        simp = [sp.dt, sp.num_substeps, sp.num_solver_iters, *sp.gravity, sp.friction, sp.erp,
                sp.contact_threshold, sp.max_contacts, sp.max_velocity, float(bool(sp.self_collision)), sp.friction_self,
                sp.warmstart, sp.spinning_friction, sp.spinning_friction_self, float(bool(sp.exact_box_contacts))]
        self.cache_dim = 2 * sp.max_contacts
        keep = []

        def D(x):
            a, p = _d(x); keep.append(a); return p

        def I(x):
            a, p = _i(x); keep.append(a); return p
        nl = len(t["parent"])
        ns = len(t["sh_link"])
        self._h = C.c_void_p(L.oc_create(
            C.c_int(nl), I(t["parent"]), I(t["jtype"]), D(t["jorg"]), D(t["com"]), D(t["mass"]),
            D(t["com_dyn"]), D(t["I_dyn"]), D(t["com_pd"]), D(t["I_pd"]),
            C.c_int(ns), I(t["sh_link"]), I(t["sh_kind"]), D(t["sh_dims"]), D(t["sh_pos"]), D(t["sh_rot"]),
            D(character.kps), D(character.kds), D(character.max_forces), D(simp),
            D(character.joint_weights), C.c_int(len(character.end_effectors)), I(character.end_effectors),
            D(character.cost_weights), C.c_double(character.height)))
        if not self._h:
            raise RuntimeError("oc_create failed")
        self.nv = L.oc_nv(self._h)
        self.nl = nl
        self.nj = len(character.movable)
        self.character = character

    def __del__(self):
        try:
            lib().oc_destroy(self._h)
        except Exception:
            pass

    # --- reference-shaped calls -------------------------------------------------------------
    def step(self, state, action, nsteps=1, cache=None):
        """nsteps x {actuate(action); stepSimulation()} -> new 132-state.  ``cache``: contact-cache row (2 * max_contacts
        values, see empty_cache) updated in place, or None to start from an empty cache and drop the result."""
        s, sp = _d(np.array(state, dtype=np.float64, copy=True))
        a, ap = _d(action)
        if cache is None:
            lib().oc_step(self._h, sp, ap, C.c_int(nsteps))
        else:
            assert cache.dtype == np.float64 and cache.flags.c_contiguous and cache.shape == (self.cache_dim,)
            lib().oc_step_cache(self._h, sp, ap, C.c_int(nsteps), cache.ctypes.data_as(C.POINTER(C.c_double)))
        return s

    def empty_cache(self, n=None):
        """Contact-cache rows with no contacts: ids -1, impulses 0."""
        c = np.zeros((n or 1, self.cache_dim))
        c[:, :self.cache_dim // 2] = -1.0
        return c if n else c[0]

    def cost(self, sim, kin):
        """-> (total, [pose, root, ee, balance, com])"""
        s, sp = _d(sim); k, kp = _d(kin)
        out = np.zeros(6)
        lib().oc_cost(self._h, sp, kp, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[0], out[1:]

    def window(self, parents, actions, target, nsteps, k0=0, k1=None, parent_cache=None, return_cache=False):
        """Samples k0..k1 of one window.  ``parent_cache`` [E, cache_dim]: the parents' contact caches (None: empty);
        with ``return_cache`` a fourth array [k1-k0, cache_dim] holds every sample's cache at the end of the window."""
        parents = np.ascontiguousarray(parents, dtype=np.float64)
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        S, E = actions.shape[0], parents.shape[0]
        assert S % E == 0
        k1 = S if k1 is None else k1
        st = np.zeros((S, parents.shape[1])); cost = np.zeros(S); info = np.zeros((S, 5))
        t, tp = _d(target)
        pc = None
        if parent_cache is not None:
            parent_cache = np.ascontiguousarray(parent_cache, dtype=np.float64)
            assert parent_cache.shape == (E, self.cache_dim)
            pc = parent_cache.ctypes.data_as(C.POINTER(C.c_double))
        oc = np.zeros((S, self.cache_dim)) if return_cache else None
        lib().oc_window_cache(self._h, _d(parents)[1], _d(actions)[1], tp, C.c_int(k0), C.c_int(k1), C.c_int(S // E),
                              C.c_int(nsteps), st.ctypes.data_as(C.POINTER(C.c_double)),
                              cost.ctypes.data_as(C.POINTER(C.c_double)), info.ctypes.data_as(C.POINTER(C.c_double)),
                              pc, oc.ctypes.data_as(C.POINTER(C.c_double)) if return_cache else None)
        if return_cache:
            return st[k0:k1], cost[k0:k1], info[k0:k1], oc[k0:k1]
        return st[k0:k1], cost[k0:k1], info[k0:k1]

    @staticmethod
    def topk(cost, k):
        c, cp = _d(cost)
        idx = np.zeros(k, dtype=np.int32)
        lib().oc_topk(cp, C.c_int(len(c)), C.c_int(k), idx.ctypes.data_as(C.POINTER(C.c_int)))
        return idx

    # --- diagnostics ------------------------------------------------------------------------
    def mass_matrix(self, state, inertia_set=0):
        M = np.zeros((self.nv, self.nv))
        lib().oc_mass_matrix(self._h, _d(state)[1], C.c_int(inertia_set), M.ctypes.data_as(C.POINTER(C.c_double)))
        return M

    def bias(self, state, inertia_set=0, gravity=True, qdd=None):
        tau = np.zeros(self.nv)
        q = _d(qdd)[1] if qdd is not None else None
        lib().oc_bias(self._h, _d(state)[1], C.c_int(inertia_set), C.c_int(int(gravity)), q,
                      tau.ctypes.data_as(C.POINTER(C.c_double)))
        return tau

    def gen_vel(self, state):
        v = np.zeros(self.nv)
        lib().oc_gen_vel(self._h, _d(state)[1], v.ctypes.data_as(C.POINTER(C.c_double)))
        return v

    def pd_torque(self, state, action):
        tau = np.zeros(self.nv)
        lib().oc_pd_torque(self._h, _d(state)[1], _d(action)[1], tau.ctypes.data_as(C.POINTER(C.c_double)))
        return tau

    def link_coms(self, state):
        pos = np.zeros((self.nl, 3)); vel = np.zeros((self.nl, 3)); com = np.zeros(6)
        lib().oc_link_coms(self._h, _d(state)[1], pos.ctypes.data_as(C.POINTER(C.c_double)),
                           vel.ctypes.data_as(C.POINTER(C.c_double)), com.ctypes.data_as(C.POINTER(C.c_double)))
        return pos, vel, com[:3], com[3:]

    def contacts(self, state):
        out = np.zeros((256, 5))
        n = lib().oc_contacts(self._h, _d(state)[1], out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[:n]

    def contacts_full(self, state):
        """rows: link, link2 (-1 = ground), point on link (3), point on link2 (3), normal link2->link (3), distance"""
        out = np.zeros((256, 12))
        n = lib().oc_contacts_full(self._h, _d(state)[1], out.ctypes.data_as(C.POINTER(C.c_double)))
        return out[:n]

    @property
    def num_pairs(self):
        return lib().oc_num_pairs(self._h)

    def last_ground_impulse(self):
        """Sum of the ground-contact impulses (world x, y, z) of the last sub-step run in this thread."""
        out = np.zeros(3)
        lib().oc_last_ground_impulse(out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def step_torque(self, state, tau, nsub_total):
        s, sp = _d(np.array(state, dtype=np.float64, copy=True))
        lib().oc_step_torque(self._h, sp, _d(tau)[1], C.c_int(nsub_total))
        return s

"""Model compiler: URDF links -> baked tables for the CUDA path (and raw tables for the oracle).

What PyBullet derives at ``loadURDF`` time (/root/reference/envs/humanoid_stable_pd.py:9-20)
is derived here once on the host:

* two rigid-body inertia sets (both [BULLET-UPSTREAM, unverified]):
  ``dyn`` -- what ``stepSimulation`` integrates: without URDF_USE_INERTIA_FROM_FILE Bullet
  replaces the URDF tensor by the inertia of the axis-aligned bounding box of the link's
  collision compound, taken in the inertial frame;
  ``pd``  -- what the stable-PD plugin's own rigid-body model uses: mass distributed over the
  link's *first* collision primitive (solid sphere / capsule / box), centred on that shape.
  Either can be switched to the other (``inertia_dyn`` / ``inertia_pd`` options).
* for the device: fixed joints folded into their parents (18 moving bodies for the SMPL
  humanoid), bodies renumbered level by level (all bodies of one tree depth are
  consecutive) so that the lanes that own one sample process one tree level per step.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .urdf import BOX, CAPSULE, SPHERE, J_FIXED, J_SPHERICAL, Link

STATE_DIM = 132          # humanoid_stable_pd.py:142-148
MAX_LEVEL_WIDTH = 4      # articulated-inertia hand-over slots per tree level in the rollout kernel (ICW)


@dataclass
class SimParams:
    """Engine settings of /root/reference/envs/humanoid_tracking_env.py:43-50, 75-76 and
    humanoid_stable_pd.py:66-76, plus Bullet defaults the reference leaves implicit
    ([BULLET-UPSTREAM, unverified])."""
    dt: float = 1.0 / 1000.0          # setTimeStep(1/fps_sim)
    num_substeps: int = 2             # numSubSteps
    num_solver_iters: int = 10        # numSolverIterations
    gravity: tuple = (0.0, 0.0, -9.8)
    friction: float = 0.8 * 0.9       # link lateral 0.8 x plane 0.9 (product rule)
    erp: float = 0.2                  # contact ERP
    contact_threshold: float = 0.02   # contactBreakingThreshold: points closer than this join the solve
    max_contacts: int = 8             # deepest-first cap per sample (device working-set bound)
    max_velocity: float = 100.0       # btMultiBody max coordinate velocity clamp
    self_collision: bool = True       # URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS (humanoid_stable_pd.py:10-13)
    friction_self: float = 0.8 * 0.8  # link lateral 0.8 x link lateral 0.8 (product rule)
    # Contact persistence (saveBullet / restoreState carry the contact cache, humanoid_tracking_env.py:178-182, 69-71):
    # a contact that was also present in the previous sub-step starts the Gauss-Seidel sweep from warmstart x its last
    # normal impulse; 0 = stateless contacts (round 1).  SURVEY.md §9.2 lists "warm-starting on (factor 0.85)"
    # [BULLET-UPSTREAM, unverified: bullet3's btMultiBodyConstraintSolver::setupMultiBodyContactConstraint has carried its
    # warm start behind a disabled branch in several releases -- set 0.0 to restate that].
    warmstart: float = 0.85
    # spinning (torsional) friction about the contact normal: humanoid_stable_pd.py:66-76 sets spinningFriction 0.3 on
    # every link; combined coefficient per contact [BULLET-UPSTREAM, unverified] = sum over the two objects of
    # spinningFriction x lateralFriction = 0.3 x 0.8 against the plane (which has none), 2 x 0.3 x 0.8 link-vs-link.
    # Only the CPU oracle implements the row (it is 0 by default: the device contact solve holds 3 rows per contact in
    # registers and sc_create rejects a non-zero value).
    spinning_friction: float = 0.0
    spinning_friction_self: float = 0.0
    # link-vs-link narrow phase: True = a box (feet, upper arms, forearms: humanoid.urdf:76,126,248,268,316,336) against a
    # sphere or capsule is the exact box-to-swept-segment distance, and of two boxes the second is stood in for by its capsule;
    # False = every box is its capsule stand-in (one closed-form segment-segment distance per pair).  Off by default: on the
    # bench workload the exact routine costs 9-10 % of the window (profiles/r2y_exact_box_cost.log) -- a lane or two per warp
    # run it while the rest of the CTA waits at the phase barrier -- for contacts that mostly occur between the tangled legs
    # of samples the search discards.  Oracle and kernel implement both; the parity tests cover both.
    exact_box_contacts: bool = False


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def _shape_halfext_in_link(s):
    """Half extents of the shape's AABB along the link axes."""
    if s.kind == SPHERE:
        return np.full(3, s.dims[0])
    if s.kind == CAPSULE:
        h = np.array([s.dims[0], s.dims[0], s.dims[0] + 0.5 * s.dims[1]])
    else:
        h = 0.5 * s.dims
    return np.abs(s.rot) @ h


def _solid_inertia(s, mass):
    """Principal inertia (shape frame) of a uniform solid primitive."""
    if s.kind == SPHERE:
        return np.full(3, 0.4 * mass * s.dims[0] ** 2)
    if s.kind == BOX:
        x, y, z = s.dims
        return mass / 12.0 * np.array([y * y + z * z, x * x + z * z, x * x + y * y])
    r, L = s.dims[0], s.dims[1]
    vc, vs = np.pi * r * r * L, 4.0 / 3.0 * np.pi * r ** 3
    mc = mass * vc / (vc + vs)
    ms = mass - mc
    iax = 0.5 * mc * r * r + 0.4 * ms * r * r
    ipe = mc * (L * L / 12.0 + r * r / 4.0) + ms * (0.4 * r * r + 0.25 * L * L + 0.375 * L * r)
    return np.array([ipe, ipe, iax])


def shape_capsule(s):
    """Self-collision stand-in of a collision primitive: a sphere-swept segment (p0, p1, radius) in the shape's link
    frame.  Spheres and capsules are exact.  A box becomes the capsule along its longest side with the same reach
    along that side and the same cross-section area (r = half the geometric mean of the two short sides): the
    link-vs-link narrow phase then is one closed-form segment-segment distance for every pair."""
    if s.kind == SPHERE:
        return s.pos.copy(), s.pos.copy(), float(s.dims[0])
    if s.kind == CAPSULE:
        a = s.rot @ np.array([0.0, 0.0, 0.5 * s.dims[1]])
        return s.pos + a, s.pos - a, float(s.dims[0])
    k = int(np.argmax(s.dims))
    short = [s.dims[i] for i in range(3) if i != k]
    r = 0.5 * float(np.sqrt(short[0] * short[1]))
    e = np.zeros(3)
    e[k] = max(0.0, 0.5 * s.dims[k] - r)
    a = s.rot @ e
    return s.pos + a, s.pos - a, r


def self_collision_pairs(links):
    """Shape pairs (i, j), i < j in URDF shape order, whose links are different and not in an ancestor relation
    (URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS, humanoid_stable_pd.py:10-13)."""
    anc = []
    for i, l in enumerate(links):
        a = set()
        p = l.parent
        while p >= 0:
            a.add(p)
            p = links[p].parent
        anc.append(a)
    sh_link = [i for i, l in enumerate(links) for _ in l.shapes]
    pairs = []
    for a in range(len(sh_link)):
        for b in range(a + 1, len(sh_link)):
            la, lb = sh_link[a], sh_link[b]
            if la != lb and la not in anc[lb] and lb not in anc[la]:
                pairs.append((a, b))
    return pairs


def link_inertia(link: Link, mode: str):
    """-> (com (3,), I (3,3) about that com in link axes) for one inertia model."""
    if mode == "aabb":
        if not link.shapes:
            return link.com.copy(), np.zeros((3, 3))
        lo = np.full(3, np.inf)
        hi = np.full(3, -np.inf)
        for s in link.shapes:
            h = _shape_halfext_in_link(s)
            lo = np.minimum(lo, s.pos - link.com - h)
            hi = np.maximum(hi, s.pos - link.com + h)
        lx, ly, lz = hi - lo
        I = link.mass / 12.0 * np.diag([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
        return link.com.copy(), I
    if mode == "first_shape":
        if not link.shapes:
            return link.com.copy(), np.zeros((3, 3))
        s = link.shapes[0]
        I = s.rot @ np.diag(_solid_inertia(s, link.mass)) @ s.rot.T
        return s.pos.copy(), I
    raise ValueError(mode)


@dataclass
class Character:
    """Host-side description shared by the oracle bindings and the device baker."""
    links: list
    kps: np.ndarray
    kds: np.ndarray
    max_forces: np.ndarray
    joint_weights: np.ndarray
    end_effectors: list
    cost_weights: np.ndarray
    height: float
    sim: SimParams = field(default_factory=SimParams)
    inertia_dyn: str = "aabb"
    inertia_pd: str = "first_shape"

    @property
    def movable(self):
        return [i for i, l in enumerate(self.links) if l.jtype == J_SPHERICAL]

    def __post_init__(self):
        nj = len(self.movable)
        for name in ("kps", "kds", "max_forces", "joint_weights"):
            a = np.asarray(getattr(self, name), dtype=np.float64)
            if a.shape != (nj,):
                raise ValueError(f"{name}: expected {nj} values, got {a.shape}")
            setattr(self, name, a)
        self.cost_weights = np.asarray(self.cost_weights, dtype=np.float64)
        if self.cost_weights.shape != (5,):
            raise ValueError("cost_weights needs 5 entries (pose, root, ee, balance, com)")
        if 7 + 4 * nj + 6 + 3 * nj != STATE_DIM:
            raise ValueError("character does not match the 132-float SamCon state layout")

    @classmethod
    def from_config(cls, cfg, sim: SimParams | None = None, **kw):
        from .urdf import load_links
        links = load_links(cfg.model_file, cfg.scale, search_dirs=(getattr(cfg, "base_dir", ""),))
        # setTimeStep(1 / cfg.fps_sim) (humanoid_tracking_env.py:25, 46); cfg key self_collision (walk.yml:113) -> loadURDF
        # flags, humanoid_stable_pd.py:10-13
        dt = 1.0 / float(cfg.fps_sim)
        if sim is None:
            sim = SimParams(dt=dt, self_collision=bool(getattr(cfg, "self_collision", True)))
        elif abs(sim.dt - dt) > 1e-12 * dt:
            raise ValueError(f"SimParams.dt = {sim.dt} contradicts cfg.fps_sim = {cfg.fps_sim}: the window length "
                             "(fps_sim // fps_act steps) and the stable-PD dt * Kd term both follow cfg.fps_sim")
        return cls(links, cfg.kps, cfg.kds, cfg.max_forces, np.asarray(cfg.joint_weights, dtype=np.float64),
                   list(cfg.end_effectors), np.asarray(cfg.cost_weights, dtype=np.float64), float(cfg.height),
                   sim, **kw)

    # ------------------------------------------------------------------ oracle tables
    def raw_tables(self):
        """Per-URDF-link arrays (fixed links kept) -- input of oracle/samcon_oracle.c."""
        L = self.links
        n = len(L)
        t = {
            "parent": np.array([l.parent for l in L], dtype=np.int32),
            "jtype": np.array([l.jtype for l in L], dtype=np.int32),
            "jorg": np.array([l.jorg for l in L], dtype=np.float64),
            "com": np.array([l.com for l in L], dtype=np.float64),
            "mass": np.array([l.mass for l in L], dtype=np.float64),
        }
        for key, mode in (("dyn", self.inertia_dyn), ("pd", self.inertia_pd)):
            ci = [link_inertia(l, mode) for l in L]
            t["com_" + key] = np.array([c for c, _ in ci], dtype=np.float64)
            t["I_" + key] = np.array([I for _, I in ci], dtype=np.float64).reshape(n, 9)
        sh = [(i, s) for i, l in enumerate(L) for s in l.shapes]
        t["sh_link"] = np.array([i for i, _ in sh], dtype=np.int32)
        t["sh_kind"] = np.array([s.kind for _, s in sh], dtype=np.int32)
        t["sh_dims"] = np.array([s.dims for _, s in sh], dtype=np.float64)
        t["sh_pos"] = np.array([s.pos for _, s in sh], dtype=np.float64)
        t["sh_rot"] = np.array([s.rot.reshape(9) for _, s in sh], dtype=np.float64)
        return t

    # ------------------------------------------------------------------ device tables
    def device_tables(self):
        """Fold fixed links, renumber by tree level and pack what the kernels read.

        Returns a dict of small numpy arrays; ``samcon_b200.pool`` flattens it into the
        ``sc_model`` blob of include/samcon_b200.h.
        """
        L = self.links
        n = len(L)
        # moving ancestor + offset of every link frame inside it
        owner = np.zeros(n, dtype=np.int64)
        off = np.zeros((n, 3))
        for i, l in enumerate(L):
            if l.jtype == J_FIXED:
                owner[i] = owner[l.parent]
                off[i] = off[l.parent] + l.jorg
            else:
                owner[i] = i
        moving = [i for i, l in enumerate(L) if l.jtype != J_FIXED]
        depth = {}
        for i in moving:
            depth[i] = 0 if L[i].parent < 0 else depth[owner[L[i].parent]] + 1
        nlev = max(depth.values()) + 1
        order = []
        level_start = [0]
        for d in range(nlev):
            lev = [i for i in moving if depth[i] == d]
            if len(lev) > MAX_LEVEL_WIDTH:
                raise ValueError(f"tree level {d} has {len(lev)} bodies; the rollout kernel handles <= {MAX_LEVEL_WIDTH}")
            order += lev
            level_start.append(len(order))
        body_of_link = {li: b for b, li in enumerate(order)}
        nb = len(order)
        movable = self.movable
        parent = np.full(nb, -1, dtype=np.int32)
        jidx = np.full(nb, -1, dtype=np.int32)
        jorg = np.zeros((nb, 3))
        kp = np.zeros(nb)
        kd = np.zeros(nb)
        maxf = np.zeros(nb)
        for b, li in enumerate(order):
            l = L[li]
            if l.parent >= 0:
                parent[b] = body_of_link[owner[l.parent]]
                jorg[b] = off[l.parent] + l.jorg
                k = movable.index(li)
                jidx[b] = k
                kp[b], kd[b], maxf[b] = self.kps[k], self.kds[k], self.max_forces[k]
        # spatial inertia about the body origin for both inertia sets: m, h = m c, I_o (xx,xy,xz,yy,yz,zz)
        inert = {}
        for key, mode in (("dyn", self.inertia_dyn), ("pd", self.inertia_pd)):
            m = np.zeros(nb)
            h = np.zeros((nb, 3))
            Io = np.zeros((nb, 3, 3))
            for i, l in enumerate(L):
                c, I = link_inertia(l, mode)
                c = c + off[i]
                b = body_of_link[owner[i]]
                m[b] += l.mass
                h[b] += l.mass * c
                Io[b] += I + l.mass * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
            inert[key] = np.concatenate([m[:, None], h, Io.reshape(nb, 9)[:, [0, 1, 2, 4, 5, 8]]], axis=1)
        # cost points: the CoM of every URDF link (getLinkStates()[0], humanoid_stable_pd.py:218-222)
        cp_body = np.array([body_of_link[owner[i]] for i in range(n)], dtype=np.int32)
        cp_off = np.array([off[i] + L[i].com for i in range(n)])
        cp_mass = np.array([l.mass for l in L])
        ee = np.array([e + 1 for e in self.end_effectors], dtype=np.int32)   # link index -1.. -> row
        # contact candidates: every primitive becomes "sphere points" (centre, radius):
        # sphere -> 1, capsule -> its 2 cap centres, box -> 8 corners with radius 0
        cand = []
        for i, l in enumerate(L):
            b = body_of_link[owner[i]]
            for s in l.shapes:
                if s.kind == SPHERE:
                    pts, r = [np.zeros(3)], s.dims[0]
                elif s.kind == CAPSULE:
                    pts, r = [np.array([0, 0, 0.5 * s.dims[1]]), np.array([0, 0, -0.5 * s.dims[1]])], s.dims[0]
                else:
                    hx, hy, hz = 0.5 * s.dims
                    pts = [np.array([sx * hx, sy * hy, sz * hz]) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
                    r = 0.0
                for p in pts:
                    cand.append((b, *(off[i] + s.pos + s.rot @ p), r))
        cand = np.array(cand, dtype=np.float64)
        # self-collision primitives (one per URDF shape, in shape order) in the frame of the moving body that carries
        # the shape's link, and the primitive pairs that may touch
        prim_body, prim, prim_box, boxes = [], [], [], []
        for i, l in enumerate(L):
            for s in l.shapes:
                p0, p1, r = shape_capsule(s)
                prim_body.append(body_of_link[owner[i]])
                prim.append([*(off[i] + p0), *(off[i] + p1), r])
                if s.kind == BOX and self.sim.exact_box_contacts:
                    # the box itself, in the frame of the moving body: centre, axes (v_body = R v_box, row major), half extents
                    prim_box.append(len(boxes))
                    boxes.append([*(off[i] + s.pos), *s.rot.reshape(9), *(0.5 * s.dims)])
                else:
                    prim_box.append(-1)
        pair = np.array(self_collision_pairs(L), dtype=np.int32).reshape(-1, 2)
        return {
            "prim_body": np.array(prim_body, dtype=np.int32), "prim": np.array(prim, dtype=np.float64), "pair": pair,
            "prim_box": np.array(prim_box, dtype=np.int32), "box": np.array(boxes, dtype=np.float64).reshape(-1, 15),
            "nb": nb, "nlev": nlev, "level_start": np.array(level_start, dtype=np.int32),
            "parent": parent, "jidx": jidx, "jorg": jorg, "kp": kp, "kd": kd, "maxf": maxf,
            "inertia_dyn": inert["dyn"], "inertia_pd": inert["pd"],
            "cp_body": cp_body, "cp_off": cp_off, "cp_mass": cp_mass, "ee": ee,
            "cand_body": cand[:, 0].astype(np.int32), "cand_pt": cand[:, 1:4], "cand_r": cand[:, 4],
            "joint_weights": self.joint_weights.copy(), "cost_weights": self.cost_weights.copy(),
            "height": self.height, "body_link": np.array(order, dtype=np.int32),
        }

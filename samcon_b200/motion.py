"""Reference-motion access: numpy restatement of the reference's ``AmassMotionData``
(/root/reference/data_loaders/amass_motion_data.py:12-109) and of the PyBullet helpers it calls
(/root/reference/utils/bullet_utils.py:9-45), plus a synthetic SMPL-shaped motion generator used where the
reference reads ``data/motion/amass_bullet.pkl`` (not shipped, /root/reference/.gitignore:123).

Motion dict format (process_amass/amass_to_bullet.py:118-121): ``{seq: {'fr': fps, 'pose': (T, 75) float64}}``
with pose = root position (3), root quaternion xyzw (4), 17 joint quaternions xyzw (68).

The three PyBullet calls are restated from bullet3's published quaternion code [BULLET-UPSTREAM, unverified]:
``getQuaternionSlerp`` = shortest-arc slerp that returns the start orientation when the two are (anti)parallel;
``getAxisAngleFromQuaternion`` = (xyz / sqrt(1 - w^2) or (1,0,0) if that is ~0, 2*acos(w)) with no
shortest-path flip; ``multiplyTransforms`` with zero positions = quaternion product.
"""
from __future__ import annotations

import math

import numpy as np

from .model import STATE_DIM

_EPS = np.finfo(np.float64).eps


def quat_mul(a, b):
    x, y, z, w = a
    X, Y, Z, W = b
    return np.array([w * X + x * W + y * Z - z * Y, w * Y - x * Z + y * W + z * X, w * Z + x * Y - y * X + z * W,
                     w * W - x * X - y * Y - z * Z], dtype=np.float64)


def quat_slerp(q0, q1, t):
    q0 = np.asarray(q0, dtype=np.float64)
    q1 = np.asarray(q1, dtype=np.float64)
    mag = math.sqrt(np.dot(q0, q0) * np.dot(q1, q1))
    prod = float(np.dot(q0, q1)) / mag
    ap = abs(prod)
    if ap < 1.0 - _EPS:
        theta = math.acos(ap)
        d = math.sin(theta)
        sign = -1.0 if prod < 0 else 1.0
        s0 = math.sin((1.0 - t) * theta) / d
        s1 = math.sin(sign * t * theta) / d
        return q0 * s0 + q1 * s1
    return q0.copy()


def axis_angle_from_quat(q):
    x, y, z, w = (float(v) for v in q)
    s2 = 1.0 - w * w
    if s2 < 10.0 * _EPS:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = np.array([x, y, z]) / math.sqrt(s2)
    return axis, 2.0 * math.acos(max(-1.0, min(1.0, w)))


def compute_lin_vel(p0, p1, dt):          # bullet_utils.py:9-15
    return (np.asarray(p1, dtype=np.float64) - np.asarray(p0, dtype=np.float64)) / dt


def compute_ang_vel(q0, q1, dt):          # bullet_utils.py:17-30, world frame (base)
    q0c = np.array([-q0[0], -q0[1], -q0[2], q0[3]])
    axis, ang = axis_angle_from_quat(quat_mul(q1, q0c))
    return axis * ang / dt


def compute_ang_vel_rel(q0, q1, dt):      # bullet_utils.py:32-45, local frame (joints)
    q0c = np.array([-q0[0], -q0[1], -q0[2], q0[3]])
    axis, ang = axis_angle_from_quat(quat_mul(q0c, q1))
    return axis * ang / dt


class AmassMotionData:
    """Same constructor and methods as the reference class; ``path`` may also be an in-memory motion dict."""

    def __init__(self, path, seq):
        if isinstance(path, dict):
            data = path[seq]
        else:
            import joblib
            data = joblib.load(path)[seq]
        self._duration = 1.0 / data["fr"]
        self._motion = np.asarray(data["pose"], dtype=np.float64)

    def numFrames(self):
        return self._motion.shape[0]

    def keyFrameDuration(self):
        return self._duration

    def getCycleTime(self):
        return self.keyFrameDuration() * (self.numFrames() - 1)

    def calcCycleCount(self, curTime, cycleTime):
        return math.floor(curTime / cycleTime)

    def computeCycleOffset(self):
        return self._motion[-1][:3] - self._motion[0][:3]

    def getSpecTimeState(self, t):
        curTime = t
        keyFrameDuration = self.keyFrameDuration()
        cycleTime = self.getCycleTime()
        cycleCount = self.calcCycleCount(curTime, cycleTime)
        frameTime = curTime - cycleCount * cycleTime
        if frameTime < 0:
            frameTime += cycleTime
        frame = int(frameTime / keyFrameDuration)
        frameNext = frame + 1
        if frameNext >= self.numFrames():
            frameNext = frame
        frameFraction = (frameTime - frame * keyFrameDuration) / keyFrameDuration
        state = self.interpolator(frameFraction, self._motion[frame], self._motion[frameNext])
        state[0:3] = state[0:3] + cycleCount * self.computeCycleOffset()
        return state

    def interpolator(self, frameFraction, frameData, frameDataNext):
        dt = self.keyFrameDuration()
        basePos = frameData[0:3] + frameFraction * (frameDataNext[0:3] - frameData[0:3])
        baseLinVel = compute_lin_vel(frameData[0:3], frameDataNext[0:3], dt)
        baseOrn = quat_slerp(frameData[3:7], frameDataNext[3:7], frameFraction)
        baseAngVel = compute_ang_vel(frameData[3:7], frameDataNext[3:7], dt)
        j0 = frameData[7:].reshape(-1, 4)
        j1 = frameDataNext[7:].reshape(-1, 4)
        rots = [quat_slerp(a, b, frameFraction) for a, b in zip(j0, j1)]
        vels = [compute_ang_vel_rel(a, b, dt) for a, b in zip(j0, j1)]
        return np.concatenate((basePos, baseOrn, np.concatenate(rots), baseLinVel, baseAngVel, np.concatenate(vels)))

    def all_window_states(self, nIter, sample_timestep):
        """Targets of windows 0..nIter as one (nIter+1, 132) array (what SamCon.sample asks for one by one)."""
        return np.stack([self.getSpecTimeState(t * sample_timestep) for t in range(nIter + 1)])


# ----------------------------------------------------------------------------------------------------------
# synthetic SMPL-shaped motions: in the loader's dict format, deterministic
# ----------------------------------------------------------------------------------------------------------
JOINTS = ["lhip", "lknee", "lankle", "rhip", "rknee", "rankle", "lowerback", "upperback", "chest", "lowerneck",
          "upperneck", "lclavicle", "lshoulder", "lelbow", "rclavicle", "rshoulder", "relbow"]


def _qx(a):
    return np.array([math.sin(0.5 * a), 0.0, 0.0, math.cos(0.5 * a)])


def _qaxis(axis, a):
    axis = np.asarray(axis, dtype=np.float64)
    return np.concatenate([math.sin(0.5 * a) * axis / np.linalg.norm(axis), [math.cos(0.5 * a)]])


def _lowest_point(links, pose, only=None):
    """Height of the lowest collision point of the character (or of the links named in `only`) for one (75,) pose with
    root z = 0."""
    from .urdf import BOX, CAPSULE, J_SPHERICAL
    from scipy.spatial.transform import Rotation as sRot
    R = [None] * len(links)
    p = [None] * len(links)
    R[0] = sRot.from_quat(pose[3:7]).as_matrix()
    p[0] = np.array([0.0, 0.0, 0.0])
    j = 0
    zmin = np.inf
    for i, l in enumerate(links):
        if i:
            Rj = np.eye(3)
            if l.jtype == J_SPHERICAL:
                Rj = sRot.from_quat(pose[7 + 4 * j:11 + 4 * j]).as_matrix()
                j += 1
            R[i] = R[l.parent] @ Rj
            p[i] = p[l.parent] + R[l.parent] @ l.jorg
        for s in (l.shapes if only is None or l.name in only else ()):
            c = p[i] + R[i] @ s.pos
            Rs = R[i] @ s.rot
            if s.kind == BOX:
                h = 0.5 * s.dims
                z = c[2] - np.abs(Rs[2]) @ h
            elif s.kind == CAPSULE:
                z = c[2] - abs(Rs[2, 2]) * 0.5 * s.dims[1] - s.dims[0]
            else:
                z = c[2] - s.dims[0]
            zmin = min(zmin, z)
    return zmin


def make_synthetic_motion(kind="walk", T=91, fr=30, seed=0, links=None, name=None):
    """-> {name: {'fr': fr, 'pose': (T,75)}}.

    gait: the planted-foot walk of make_gait_motion (what solves and bench.py use).
    walk: upright SMPL-like gait, root travels along the facing direction (-Y world: the character's links hang
    along local -Y and toes point to local +Z, so the AMASS "upright" root rotation is +90 deg about X), hips
    +-0.45 rad at 1 Hz in antiphase, knees 0..0.8 rad, ankles compensating, shoulders swinging +-0.3 rad, small
    seeded per-sequence variation.  cartwheel: whole-body rotation of 2*pi about the forward axis over 1.5 s with
    arms raised so that hands reach the ground, legs spread.
    Root height is set frame by frame so that the lowest collision point touches the ground.
    """
    if kind == "gait":
        return make_gait_motion(T=T, fr=fr, seed=seed, links=links, name=name)
    if kind == "touchdown":
        return make_touchdown_motion(T=T, fr=fr, seed=seed, links=links, name=name)
    if links is None:
        from .urdf import load_links
        links = load_links(None)
    rng = np.random.default_rng(seed)
    amp = 1.0 + 0.1 * rng.standard_normal(6)
    phase0 = rng.uniform(0, 2 * np.pi)
    name = name or f"synthetic_{kind}_{seed}"
    up = _qx(0.5 * np.pi)
    pose = np.zeros((T, 75))
    for f in range(T):
        t = f / fr
        q = {n: np.array([0.0, 0.0, 0.0, 1.0]) for n in JOINTS}
        if kind == "walk":
            ph = 2 * np.pi * 1.0 * t + phase0
            hip = 0.45 * amp[0]
            # flexion (leg forward, local +Z) is a negative rotation about local X
            q["lhip"] = _qx(-hip * math.sin(ph))
            q["rhip"] = _qx(hip * math.sin(ph))
            q["lknee"] = _qx(0.8 * amp[1] * max(0.0, math.sin(ph - 0.5 * np.pi)) ** 2 + 0.05)
            q["rknee"] = _qx(0.8 * amp[1] * max(0.0, math.sin(ph + 0.5 * np.pi)) ** 2 + 0.05)
            q["lankle"] = _qx(0.15 * amp[2] * math.sin(ph + 0.3))
            q["rankle"] = _qx(-0.15 * amp[2] * math.sin(ph + 0.3))
            q["lshoulder"] = quat_mul(_qaxis([0, 0, 1], -1.2), _qx(0.3 * amp[3] * math.sin(ph)))
            q["rshoulder"] = quat_mul(_qaxis([0, 0, 1], 1.2), _qx(-0.3 * amp[3] * math.sin(ph)))
            q["lelbow"] = _qaxis([0, 1, 0], -0.3)
            q["relbow"] = _qaxis([0, 1, 0], 0.3)
            q["upperback"] = _qaxis([0, 1, 0], 0.08 * amp[4] * math.sin(ph))
            root_q = quat_mul(_qaxis([0, 0, 1], 0.03 * amp[5] * math.sin(ph)), up)
            root_xy = np.array([0.0, -1.0 * t])
        elif kind == "cartwheel":
            dur = 1.5
            s = min(1.0, max(0.0, (t - 0.5) / dur))
            s = 0.5 - 0.5 * math.cos(np.pi * s)                       # ease in/out
            # spin about the forward axis (world -Y) after standing for 0.5 s
            root_q = quat_mul(_qaxis([0, -1, 0], 2 * np.pi * s), up)
            spread = 0.5 * math.sin(np.pi * s) * amp[0]
            q["lhip"] = _qaxis([0, 0, 1], spread)
            q["rhip"] = _qaxis([0, 0, 1], -spread)
            raise_ = 1.4 * math.sin(np.pi * s) * amp[1]
            q["lshoulder"] = _qaxis([0, 0, 1], -1.2 + raise_ + 1.2 * math.sin(np.pi * s))
            q["rshoulder"] = _qaxis([0, 0, 1], 1.2 - raise_ - 1.2 * math.sin(np.pi * s))
            root_xy = np.array([-1.2 * s, 0.0])
        else:
            raise ValueError(kind)
        pose[f, 3:7] = root_q / np.linalg.norm(root_q)
        for k, n in enumerate(JOINTS):
            pose[f, 7 + 4 * k:11 + 4 * k] = q[n]
        pose[f, 0:2] = root_xy
        pose[f, 2] = -_lowest_point(links, pose[f]) + 0.001
    # keep quaternion signs continuous so finite-difference velocities stay small (see module docstring)
    for f in range(1, T):
        for k in range(18):
            a, b = pose[f - 1, 3 + 4 * k:7 + 4 * k], pose[f, 3 + 4 * k:7 + 4 * k]
            if np.dot(a, b) < 0:
                pose[f, 3 + 4 * k:7 + 4 * k] = -b
    return {name: {"fr": fr, "pose": pose}}


def check_state(state):
    s = np.asarray(state)
    if s.shape[-1] != STATE_DIM:
        raise ValueError(f"expected {STATE_DIM}-float states")
    return s


# ----------------------------------------------------------------------------------------------------------
# planted-foot gait: a kinematically consistent synthetic walk (stance feet do not slide, pelvis sways over the
# stance foot, the motion starts from standing).  The sinusoid walk above is kept for the parity tests that use
# it as a source of poses; a SamCon *solve* needs a motion a physical character can follow.
# ----------------------------------------------------------------------------------------------------------
def _rot_x(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _rot_z(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def _mat_to_quat(R):
    from scipy.spatial.transform import Rotation as sRot
    return sRot.from_matrix(R).as_quat()


def _leg_ik(hip_org, knee_org, ankle_org, want, x0):
    """(hip roll about z, hip flexion about x, knee flexion about x) such that the ankle joint sits at `want`
    (pelvis frame).  Newton iterations with a finite-difference Jacobian, started from the previous frame."""
    def ankle(x):
        Rh = _rot_z(x[0]) @ _rot_x(x[1])
        return hip_org + Rh @ (knee_org + _rot_x(x[2]) @ ankle_org)
    x = np.array(x0, dtype=np.float64)
    for _ in range(30):
        e = ankle(x) - want
        if np.linalg.norm(e) < 1e-10:
            break
        J = np.zeros((3, 3))
        for k in range(3):
            d = np.zeros(3); d[k] = 1e-6
            J[:, k] = (ankle(x + d) - ankle(x - d)) / 2e-6
        x = x - np.linalg.solve(J + 1e-9 * np.eye(3), e)
        x[2] = max(x[2], 0.02)            # the knee only bends one way
    return x


def make_gait_motion(T=91, fr=30, seed=0, links=None, name=None, step_time=0.6, step_length=0.36, double_support=0.25,
                     pelvis_height=0.93, half_width=0.09, sway=0.04, lift=0.06, ramp=0.3):
    """-> {name: {'fr': fr, 'pose': (T,75)}}: a walk built from footholds.

    The pelvis advances along the facing direction (-Y world) with speed step_length/step_time after a smooth start
    from standing, sways towards the stance side, and each foot alternates between standing flat on its foothold
    and a cycloidal swing to the next one; hip and knee angles come from inverse kinematics on the character's own
    link offsets, the ankle keeps the sole level.  Arms swing against the legs.  `seed` varies step length, cadence
    and sway by a few percent so that different seeds are different sequences.

    Defaults (0.6 m/s): chosen inside the region where independent walk.yml-size searches (nSample 2000 / nSave 200) all
    keep their best elite cost below ~2.5 for 30 windows on the B200 (profiles/r2a_gait_sweep.log; 0.4 m steps every
    0.5 s still track at <= 4.1, 0.5 m steps or 0.4 s steps do not).
    """
    if links is None:
        from .urdf import load_links
        links = load_links(None)
    rng = np.random.default_rng(seed)
    var = 1.0 + (0.05 * rng.standard_normal(3) if seed else np.zeros(3))
    Ts, Ls, sway = step_time * var[0], step_length * var[1], sway * var[2]
    v = Ls / Ts
    name = name or f"synthetic_gait_{seed}"
    idx = {l.name: i for i, l in enumerate(links)}
    jorg = {n: links[idx[n]].jorg for n in idx}
    sole = 0.09 + 0.001                                       # ankle joint above the ground when the sole is level

    def progress(t):                                          # pelvis travel, velocity v (1 - exp(-t/ramp))
        return v * (t - ramp * (1.0 - math.exp(-t / ramp)))

    def gain(t):
        return 1.0 - math.exp(-t / ramp)

    # footholds: foot f (0 left, 1 right) swings during steps k = f, f+2, ...; it lands where the pelvis will be at
    # the middle of the stance that follows
    def stance_mid(k):                                        # foot that swings in step k stands from (k+1) Ts
        return (k + 1) * Ts + 0.5 * (1.0 + double_support) * Ts

    def foot(f, t):
        """-> (forward position, height above the level stance) of foot f at time t"""
        k = int(math.floor(t / Ts))
        # last step (<= k) in which this foot swings
        ks = k if k % 2 == f else k - 1
        prev = progress(stance_mid(ks - 2)) if ks - 2 >= -1 and ks >= 2 else 0.0
        if ks < 0:
            return 0.0, 0.0
        nxt = progress(stance_mid(ks))
        if ks == k:                                           # swinging (after the double-support part of the step)
            s = (t - k * Ts) / Ts
            s = (s - double_support) / (1.0 - double_support)
            if s <= 0:
                return prev, 0.0
            c = s - math.sin(2 * math.pi * s) / (2 * math.pi)
            return prev + (nxt - prev) * c, lift * math.sin(math.pi * s) ** 2
        return nxt, 0.0

    up = _qx(0.5 * np.pi)
    pose = np.zeros((T, 75))
    x_prev = [np.array([0.0, -0.3, 0.6]), np.array([0.0, -0.3, 0.6])]
    for fidx in range(T):
        t = fidx / fr
        px = -sway * gain(t) * math.sin(math.pi * (t - 0.5 * double_support * Ts) / Ts)   # towards the stance side
        pz = progress(t)
        q = {n: np.array([0.0, 0.0, 0.0, 1.0]) for n in JOINTS}
        fz = []
        for f, (hip, knee, ank) in enumerate((("lhip", "lknee", "lankle"), ("rhip", "rknee", "rankle"))):
            z, h = foot(f, t)
            fz.append(z - pz)
            side = 1.0 if f == 0 else -1.0
            want = np.array([side * half_width - px, sole + h - pelvis_height, z - pz])    # pelvis frame: x left, y up, z fwd
            x = _leg_ik(jorg[hip], jorg[knee], jorg[ank], want, x_prev[f])
            x_prev[f] = x
            Rh = _rot_z(x[0]) @ _rot_x(x[1])
            Rk = _rot_x(x[2])
            q[hip] = _mat_to_quat(Rh)
            q[knee] = _mat_to_quat(Rk)
            q[ank] = _mat_to_quat((Rh @ Rk).T)              # sole level, toes forward
        swing = 0.3 / 0.25
        q["lshoulder"] = quat_mul(_qaxis([0, 0, 1], -1.2), _qx(swing * fz[1]))
        q["rshoulder"] = quat_mul(_qaxis([0, 0, 1], 1.2), _qx(swing * fz[0]))
        q["lelbow"] = _qaxis([0, 1, 0], -0.3)
        q["relbow"] = _qaxis([0, 1, 0], 0.3)
        pose[fidx, 0:3] = [px, -pz, pelvis_height]
        pose[fidx, 3:7] = up
        for k, n in enumerate(JOINTS):
            pose[fidx, 7 + 4 * k:11 + 4 * k] = q[n]
    for fidx in range(1, T):
        for k in range(18):
            a, b = pose[fidx - 1, 3 + 4 * k:7 + 4 * k], pose[fidx, 3 + 4 * k:7 + 4 * k]
            if np.dot(a, b) < 0:
                pose[fidx, 3 + 4 * k:7 + 4 * k] = -b
    return {name: {"fr": fr, "pose": pose}}


# ----------------------------------------------------------------------------------------------------------
# "touchdown": a contact-rich motion a physical character can follow -- stand, fold forward until both hands rest on
# the ground in front of the planted feet (hand AND foot ground contact, what cartwheel.yml's cost weights and noise
# table are made for), hold, stand up again.  The kinematic cartwheel above rotates the whole body like a rigid wheel
# from a standing start; no search follows it (200 000 samples per window end up lying on the ground by window 15,
# tools/cartwheel_probe.py), so solves and bench.py's config 3 run on this motion and the cartwheel stays what it is good
# for: a source of hand-stand states for the parity tests.
# ----------------------------------------------------------------------------------------------------------
def make_touchdown_motion(T=100, fr=30, seed=0, links=None, name=None, fold_time=1.2, hold_time=0.6, start=0.4,
                          half_width=0.10, hips_back=0.85, shin=0.05, spine=0.2, hand_clearance=-0.005):
    """-> {name: {'fr': fr, 'pose': (T,75)}}.  Sagittal and symmetric: feet stay flat on their footholds (the root pose
    follows from the leg chain), the pelvis pitch at full fold is chosen so that the hands' spheres press
    `-hand_clearance` m into the ground, arms hang along the world's vertical throughout."""
    if links is None:
        from .urdf import load_links
        links = load_links(None)
    rng = np.random.default_rng(seed)
    var = 1.0 + (0.05 * rng.standard_normal(3) if seed else np.zeros(3))
    fold_time, hips_back, spine = fold_time * var[0], hips_back * var[1], spine * var[2]
    name = name or f"synthetic_touchdown_{seed}"
    idx = {l.name: i for i, l in enumerate(links)}
    jorg = {n: links[idx[n]].jorg for n in idx}
    sole = 0.09 + 0.001
    up_R = _rot_x(0.5 * np.pi)

    def frame(s, phi_max):
        """joint quaternions + root pose (upright frame: x left, y up, z forward) at fold fraction s in [0, 1]"""
        phi = phi_max * s                                    # pelvis pitch (forward)
        psi_t = -hips_back * s                               # thigh's pitch in the world: hips go back (and down) so that the
        a_h = psi_t - phi                                    # centre of mass stays over the feet while the torso comes forward
        a_k = shin * s - psi_t                               # knees bend; the shins stay close to vertical
        a_a = -(phi + a_h + a_k)                             # ankles keep the soles level
        q = {n: np.array([0.0, 0.0, 0.0, 1.0]) for n in JOINTS}
        for side, (hip, kn, ank) in ((1.0, ("lhip", "lknee", "lankle")), (-1.0, ("rhip", "rknee", "rankle"))):
            q[hip], q[kn], q[ank] = _qx(a_h), _qx(a_k), _qx(a_a)
        q["lowerback"] = _qx(spine * s)
        q["upperback"] = _qx(spine * s)
        beta = phi + 2 * spine * s                           # pitch of the chest: arms hang along the world's vertical
        q["lshoulder"] = _mat_to_quat(_rot_x(-beta) @ _rot_z(-0.5 * np.pi))
        q["rshoulder"] = _mat_to_quat(_rot_x(-beta) @ _rot_z(0.5 * np.pi))
        # root pose from the planted left foot upwards (the right leg mirrors it)
        R_root = _rot_x(phi)
        R_th = R_root @ _rot_x(a_h)
        R_sh = R_th @ _rot_x(a_k)
        ankle = np.array([half_width, sole, 0.0])
        knee_p = ankle - R_sh @ jorg["lankle"]
        hip_p = knee_p - R_th @ jorg["lknee"]
        root = hip_p - R_root @ jorg["lhip"]
        root[0] = 0.0
        return q, R_root, root

    def lowest_hand(phi_max):
        q, R_root, root = frame(1.0, phi_max)
        pose = _pack_pose(q, R_root, root, up_R)
        return _lowest_point(links, pose, only=("lwrist", "rwrist")) + pose[2]

    lo, hi = 0.8, 2.4                                        # pelvis pitch at full fold: hands reach the ground
    for _ in range(40):
        mid = 0.5 * (lo + hi)
        if lowest_hand(mid) > hand_clearance:
            lo = mid
        else:
            hi = mid
    phi_max = 0.5 * (lo + hi)
    pose = np.zeros((T, 75))
    for f in range(T):
        t = f / fr
        u = (t - start) / fold_time
        if u <= 0:
            s = 0.0
        elif u < 1:
            s = 0.5 - 0.5 * math.cos(math.pi * u)
        elif t < start + fold_time + hold_time:
            s = 1.0
        else:
            u2 = (t - start - fold_time - hold_time) / fold_time
            s = 0.5 + 0.5 * math.cos(math.pi * min(1.0, u2))
        q, R_root, root = frame(s, phi_max)
        pose[f] = _pack_pose(q, R_root, root, up_R)
    for f in range(1, T):
        for k in range(18):
            a, b = pose[f - 1, 3 + 4 * k:7 + 4 * k], pose[f, 3 + 4 * k:7 + 4 * k]
            if np.dot(a, b) < 0:
                pose[f, 3 + 4 * k:7 + 4 * k] = -b
    return {name: {"fr": fr, "pose": pose}}


def _pack_pose(q, R_root, root, up_R):
    """(joint quats by name, root rotation / position in the upright frame) -> one (75,) pose row in world coordinates."""
    pose = np.zeros(75)
    pose[0:3] = up_R @ root
    pose[3:7] = _mat_to_quat(up_R @ R_root)
    for k, n in enumerate(JOINTS):
        pose[7 + 4 * k:11 + 4 * k] = q[n]
    return pose

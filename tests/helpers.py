"""Seeded humanoid states / actions shared by the parity tests."""
import numpy as np


def qmul(a, b):
    x, y, z, w = a
    X, Y, Z, W = b
    return np.array([w * X + x * W + y * Z - z * Y, w * Y - x * Z + y * W + z * X, w * Z + x * Y - y * X + z * W,
                     w * W - x * X - y * Y - z * Z])


def rand_quat(rng, amp=0.3):
    """Rotation by a rotation vector drawn uniformly from [-amp, amp]^3 (bounded, unlike a Gaussian quaternion)."""
    if amp >= 3.0:
        q = rng.normal(size=4)
        return q / np.linalg.norm(q)
    v = rng.uniform(-amp, amp, size=3)
    a = np.linalg.norm(v)
    if a < 1e-12:
        return np.array([0.0, 0.0, 0.0, 1.0])
    return np.concatenate([np.sin(0.5 * a) * v / a, [np.cos(0.5 * a)]])


def rand_state(rng, vel=1.0, z=1.5, amp=0.3):
    """Arbitrary airborne configuration."""
    s = np.zeros(132)
    s[0:3] = [0.1, 0.2, z]
    s[3:7] = rand_quat(rng, 4.0)
    for k in range(17):
        s[7 + 4 * k:11 + 4 * k] = rand_quat(rng, amp)
    s[75:132] = rng.normal(size=57) * vel
    return s


def near_action(state, rng, amp):
    """PD target = the state's joint rotations perturbed by ~amp rad."""
    a = np.array(state[7:75]).reshape(17, 4).copy()
    for k in range(17):
        a[k] = qmul(a[k], rand_quat(rng, amp))
    return a.reshape(-1)


def stand(rng, amp=0.05, vel=0.1, z=0.99):
    """Upright humanoid (links hang along -Y, root rotated +90 deg about X) near the ground."""
    s = np.zeros(132)
    s[3:7] = [np.sin(np.pi / 4), 0, 0, np.cos(np.pi / 4)]
    for k in range(17):
        s[7 + 4 * k + 3] = 1
    s[7:75] = near_action(s, rng, amp)
    s[0:3] = [rng.normal() * 0.1, rng.normal() * 0.1, z]
    s[75:132] = rng.normal(size=57) * vel
    return s


def joint_rms(a, b):
    """RMS joint-angle difference (rad) between two batches of 132-states, per sample."""
    qa = np.asarray(a, dtype=np.float64)[:, 3:75].reshape(len(a), 18, 4)
    qb = np.asarray(b, dtype=np.float64)[:, 3:75].reshape(len(b), 18, 4)
    d = np.abs((qa * qb).sum(-1)) / (np.linalg.norm(qa, axis=-1) * np.linalg.norm(qb, axis=-1))
    ang = 2 * np.arccos(np.clip(d, 0, 1))
    return np.sqrt((ang ** 2).mean(-1))


def qz(a):
    return np.array([0.0, 0.0, np.sin(0.5 * a), np.cos(0.5 * a)])


def legs_crossed(rng, inward=0.2, z=1.5, vel=0.0):
    """Upright humanoid whose hips are rotated `inward` rad towards each other (thighs / shins / feet touch or
    interpenetrate from ~0.1 rad on): the seeded source of link-vs-link contacts for the self-collision tests."""
    s = stand(rng, 0.0, vel, z)
    s[0:2] = 0
    s[7:11] = qz(-inward)              # lhip (joint 0)
    s[7 + 12:11 + 12] = qz(inward)     # rhip (joint 3)
    return s

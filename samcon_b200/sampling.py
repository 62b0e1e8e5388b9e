"""Host restatement of ``SamCon.get_batch_action`` (/root/reference/samcon/core/samcon.py:101-142).

Same arithmetic and the same consumption of the legacy MT19937 stream (``np.random.seed(cfg.seed)`` at
samcon.py:57, ``np.random.random`` at :115, ``np.random.shuffle`` at :140), so that with equal seeds the
actions equal the reference's bit for bit (tests/test_sampling.py pins this against the reference function
itself and against committed golden vectors).  ``draw_noise`` exposes the raw draws + permutation so the
same batch can be produced on the device by ``sc_sample_gen``.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as sRot


def draw_noise(nSample, values, noise_type="uniform", rng=None):
    """-> (noise (nSample, len(values)), perm (nSample,)).

    uniform: the raw [0,1) draws (the reference then maps them to [-v, v)); gaussian: N(0, diag(values)).
    perm is the row order np.random.shuffle applies afterwards: batch[i] <- batch[perm[i]].
    """
    rng = rng if rng is not None else np.random
    values = np.asarray(values, dtype=np.float64)
    n = len(values)
    if noise_type == "gaussian":
        noise = rng.multivariate_normal(mean=np.zeros(n), cov=np.diag(values), size=(nSample))
    elif noise_type == "uniform":
        noise = rng.random_sample((nSample, n))
    else:
        raise NotImplementedError
    perm = np.arange(nSample)
    rng.shuffle(perm)
    return noise, perm


def actions_from_noise(state, noise, values, noise_type="uniform", perm=None):
    values = np.asarray(values, dtype=np.float64)
    noise = np.array(noise, dtype=np.float64, copy=True)
    if noise_type == "uniform":
        for i in range(len(values)):            # samcon.py:116-118
            noise[:, i] *= 2 * values[i]
            noise[:, i] += -1 * values[i]
    state = np.asarray(state, dtype=np.float64)
    nj4 = (len(state) - 13) // 7 * 4
    state_euler = sRot.from_quat(state[7:7 + nj4].reshape(-1, 4)).as_euler("XYZ", degrees=False).reshape(-1)
    batch = np.zeros((noise.shape[0], len(state_euler)))
    batch[:] = state_euler
    batch += noise
    while np.any(batch > np.pi):
        batch[batch > np.pi] -= 2 * np.pi
    while np.any(batch < -np.pi):
        batch[batch < -np.pi] += 2 * np.pi
    quat = sRot.from_euler("XYZ", batch.reshape(-1, 3), degrees=False).as_quat().reshape(noise.shape[0], -1)
    if perm is not None:
        quat = quat[np.asarray(perm)]
    return quat


def get_batch_action(state, nSample, values, noise_type="uniform", rng=None):
    """Drop-in for SamCon.get_batch_action: (nSample, 68) float64 xyzw actions, rows shuffled."""
    noise, perm = draw_noise(nSample, values, noise_type, rng)
    return actions_from_noise(state, noise, values, noise_type, perm)

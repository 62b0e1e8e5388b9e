"""§8 f2 on the device: sc_motion_states (slerp + finite-difference target states for every (sequence, window) pair in one
launch) against the host restatement of the reference loader (samcon_b200/motion.py:AmassMotionData.getSpecTimeState <-
/root/reference/data_loaders/amass_motion_data.py:48-109), including frame-boundary times, times past the end of the cycle
(cycle offset) and sequences of different length and frame rate; and one sc_sample_gen launch for many sequences."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _loaders():
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    out = []
    for kind, T, fr, seed in (("gait", 100, 30, 0), ("walk", 46, 30, 3), ("cartwheel", 76, 30, 0), ("gait", 61, 20, 2)):
        m = make_synthetic_motion(kind, T=T, fr=fr, seed=seed)
        (name, _), = m.items()
        out.append(AmassMotionData(m, name))
    return out


def test_motion_states_match_host_loader(pool):
    lds = _loaders()
    times = np.concatenate([np.arange(0, 41) * 0.1, [0.0333, 1.0 / 30.0, 2.0 / 30.0 + 1e-9, 0.9999, 1.5, 3.3, 4.75, 7.01]])
    dev = pool.motion_states(lds, times).cpu().numpy().astype(np.float64)
    assert dev.shape == (len(lds), len(times), 132)
    for q, ld in enumerate(lds):
        ref = np.stack([ld.getSpecTimeState(float(t)) for t in times])
        # positions / quaternions: float32 storage of the poses; velocities: finite differences of float32 poses over 1/fr
        np.testing.assert_allclose(dev[q, :, :75], ref[:, :75], atol=3e-6)
        np.testing.assert_allclose(dev[q, :, 75:], ref[:, 75:], atol=2e-4, rtol=1e-4)
    # the frame the reference picks at exact multiples of the key-frame duration
    ld = lds[0]
    for f in (1, 7, 30, 60):
        t = f * ld.keyFrameDuration()
        ref = ld.getSpecTimeState(t)
        d = pool.motion_states([ld], [t]).cpu().numpy()[0, 0]
        np.testing.assert_allclose(d[75:], ref[75:], atol=2e-4, rtol=1e-4)


def test_batched_solver_uses_device_targets(pool, walk_cfg):
    """BatchedSamCon's targets for all Q x (nIter + 1) windows come from sc_motion_states and equal the host loader's."""
    from samcon_b200.batched import BatchedSamCon
    lds = _loaders()[:2]
    b = BatchedSamCon(walk_cfg, lds, pool, nIter=5, nSample=20, nSave=2)
    for q, ld in enumerate(lds):
        ref = np.stack([ld.getSpecTimeState(0.1 * t) for t in range(6)])
        np.testing.assert_allclose(b.targets_host[q][:, :75], ref[:, :75], atol=3e-6)
        np.testing.assert_allclose(b.targets_host[q][:, 75:], ref[:, 75:], atol=2e-4, rtol=1e-4)


def test_sample_gen_for_many_sequences_in_one_launch(pool, walk_cfg):
    """Rows [q S, (q+1) S) of a Q-sequence launch are what sequence q draws alone with seed + q; host noise and permutations
    are applied per sequence."""
    import torch
    from samcon_b200.sampling import draw_noise
    lds = _loaders()[:3]
    tg = pool.motion_states(lds, [0.3])[:, 0, :].contiguous()
    lim = torch.tensor(walk_cfg.values, dtype=torch.float32, device="cuda")
    S = 50
    l0 = pool.launch_count
    both = pool.sample_gen(tg, lim, S, seed=11, window=4)
    assert pool.launch_count == l0 + 1 and both.shape == (3 * S, 68)
    for q in range(3):
        alone = pool.sample_gen(tg[q], lim, S, seed=11 + q, window=4)
        assert torch.equal(alone, both[q * S:(q + 1) * S])
    rs = np.random.RandomState(5)
    draws = [draw_noise(S, walk_cfg.values, "uniform", rs) for _ in range(3)]
    noise = np.concatenate([d[0] for d in draws]).astype(np.float32)
    perm = np.concatenate([d[1] for d in draws]).astype(np.int32)
    both = pool.sample_gen(tg, lim, S, noise=noise, perm=perm)
    for q in range(3):
        alone = pool.sample_gen(tg[q], lim, S, noise=draws[q][0].astype(np.float32), perm=draws[q][1].astype(np.int32))
        assert torch.equal(alone, both[q * S:(q + 1) * S])

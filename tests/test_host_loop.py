"""Host side of the path: motion loader restatement, Trajectory, the SamCon loop (on an oracle-backed pool)."""
import numpy as np
import pytest

from oracle_pool import OraclePool
from samcon_b200.config import Config
from samcon_b200.motion import AmassMotionData, make_synthetic_motion, quat_slerp
from samcon_b200.samcon import SamCon, UniqueId
from samcon_b200.trajectory import Trajectory


@pytest.fixture(scope="module")
def motion():
    return make_synthetic_motion("walk", T=31, seed=0)


def test_config_keys():
    cfg = Config("walk")
    assert (cfg.fps_sim, cfg.fps_act, cfg.seed, cfg.nIter, cfg.nSample, cfg.nSave) == (1000, 10, 1, 30, 2000, 200)
    assert cfg.values.shape == (51,) and cfg.kps.shape == (17,) and cfg.char_info["max_forces"][0] == 300.0
    assert cfg.end_effectors == [14, 18, 2, 5] and cfg.cost_weights == [0, 5, 20, 100, 50]
    c2 = Config("cartwheel")
    assert c2.nIter == 50 and c2.cost_weights == [0, 10, 60, 30, 10]
    with pytest.raises(FileNotFoundError):
        Config("nope")


def test_synthetic_motion_shape_and_ground(motion, character):
    (name, d), = motion.items()
    assert d["fr"] == 30 and d["pose"].shape == (31, 75)
    q = d["pose"][:, 3:].reshape(31, 18, 4)
    assert np.allclose(np.linalg.norm(q, axis=-1), 1.0)
    assert 0.9 < d["pose"][0, 2] < 1.05                           # SMPL-like root height
    assert (np.diff(d["pose"][:, 1]) < 0).all()                    # walks along the facing direction (-Y)


def test_motion_loader_interpolation(motion):
    (name, d), = motion.items()
    m = AmassMotionData(motion, name)
    assert m.numFrames() == 31 and np.isclose(m.getCycleTime(), 1.0)
    s0 = m.getSpecTimeState(0.0)
    assert s0.shape == (132,) and np.allclose(s0[:75], d["pose"][0])
    # root linear velocity = finite difference of key frames (bullet_utils.py:9-15)
    assert np.allclose(s0[75:78], (d["pose"][1, :3] - d["pose"][0, :3]) * 30)
    # halfway between frames: slerp midpoint and lerp of position
    sh = m.getSpecTimeState(0.5 / 30)
    assert np.allclose(sh[:3], 0.5 * (d["pose"][0, :3] + d["pose"][1, :3]))
    assert np.allclose(sh[3:7], quat_slerp(d["pose"][0, 3:7], d["pose"][1, 3:7], 0.5))
    # wrap-around adds the cycle offset (amass_motion_data.py:72-76)
    s_wrap = m.getSpecTimeState(1.0 + 0.1)
    s_in = m.getSpecTimeState(0.1)
    assert np.allclose(s_wrap[:3] - s_in[:3], d["pose"][-1, :3] - d["pose"][0, :3])
    assert np.allclose(s_wrap[3:], s_in[3:])
    # joint angular velocity reproduces the frame-to-frame rotation
    from scipy.spatial.transform import Rotation as R
    j = 0
    w = s0[81:84]
    r = R.from_quat(d["pose"][0, 7:11]).inv() * R.from_quat(d["pose"][1, 7:11])
    assert np.allclose(w / 30, r.as_rotvec(), atol=1e-9)


def test_unique_id_and_trajectory_paths():
    u = UniqueId()
    assert list(u.get(1)) == [0] and list(u.get(3)) == [1, 2, 3]
    nIter, S, E = 3, 6, 2
    tr = Trajectory(nIter, S, E)
    for _ in range(S):
        tr.push(0, [-1, 0, np.zeros(132), None, None, 0.0, None])
    tr.select(0)
    rng = np.random.default_rng(0)
    uid = 1
    for t in range(1, nIter + 1):
        prev = tr.get_curr_uid_elite(t - 1)
        cost = rng.random(S)
        for k in range(S):
            tr.push(t, [prev[k // 3], uid, np.full(132, t), np.full(132, uid), np.full(68, uid), cost[k], list(cost[k] * np.ones(5))])
            uid += 1
        tr.select(t, "greedy")
        assert list(tr.get_elite_inds(t)) == list(np.argsort(cost)[:E])
    out, info = tr.build_results("seq")
    assert set(out["seq"]) == {"path_0", "path_1"}
    p0 = out["seq"]["path_0"]
    assert p0["simulated_state"].shape == (3, 132) and p0["action"].shape == (3, 68) and p0["cost"].shape == (3, 1)
    assert p0["total_cost"] <= out["seq"]["path_1"]["total_cost"]
    # the path is a chain: each elite's parent is the previous step's uid
    uids = p0["simulated_state"][:, 0].astype(int)
    for a, b in zip(uids[:-1], uids[1:]):
        assert tr._parent_of[b] == a
    assert info["total_cost"].shape == (3, 6) and info["elite_inds"].shape == (3, 2) and info["best_path_my_sys"].shape == (3,)
    with pytest.raises(NotImplementedError):
        tr.select(1, "diversity")


def test_samcon_loop_on_oracle_pool(oracle, motion):
    (name, _), = motion.items()
    cfg = Config("walk")
    cfg.motion_seq = name
    loader = AmassMotionData(motion, name)
    sc = SamCon(cfg, loader, num_processes=8, nIter=2, nSample=20, nSave=4, pool=OraclePool(oracle))
    out, info = sc.sample(save=False)
    assert len(out[name]) == 4
    p0 = out[name]["path_0"]
    assert p0["simulated_state"].shape == (2, 132) and np.isfinite(p0["total_cost"])
    assert np.allclose(p0["t0_state"], loader.getSpecTimeState(0.0))
    assert np.allclose(p0["target_state"][1], loader.getSpecTimeState(0.2))
    assert info["total_cost"].shape == (2, 20)
    # elites are the cheapest samples, ascending
    for t in range(2):
        e = info["elite_inds"][t]
        assert list(e) == list(np.lexsort((np.arange(20), info["total_cost"][t]))[:4])
    # same seed -> same search (np.random.seed(cfg.seed) semantics)
    sc2 = SamCon(cfg, loader, 1, 2, 20, 4, pool=OraclePool(oracle))
    out2, _ = sc2.sample(save=False)
    assert np.array_equal(out2[name]["path_0"]["action"], p0["action"])
    with pytest.raises(Exception, match="multiples of nSave"):
        SamCon(cfg, loader, 1, 2, 21, 4, pool=OraclePool(oracle))


def test_batched_multi_sequence_equals_single_runs(oracle):
    """BASELINE config 5 (many sequences solved concurrently): one batched run over Q sequences must give every
    sequence the path it gets when solved alone -- per-sequence targets, parents, segmented elite selection and the
    back-trace are independent per sequence."""
    from oracle_pool import OracleDevicePool
    from samcon_b200.batched import BatchedSamCon
    cfg = Config("walk")
    pool = OracleDevicePool(oracle, cfg)
    loaders = []
    for seed in (0, 3):
        m = make_synthetic_motion("walk", T=16, seed=seed)
        (name, _), = m.items()
        loaders.append(AmassMotionData(m, name))
    kw = dict(nIter=2, nSample=12, nSave=3)
    both = BatchedSamCon(cfg, loaders, pool, device_noise=False, **kw).sample()
    assert len(both) == 2 and both[0]["simulated_state"].shape == (2, 132) and both[0]["action"].shape == (2, 68)
    for q, ld in enumerate(loaders):
        alone = BatchedSamCon(cfg, [ld], pool, device_noise=False, **kw).sample()[0]
        for key in ("simulated_state", "action", "cost", "elite_cost", "target_state"):
            np.testing.assert_array_equal(alone[key], both[q][key])
        assert alone["total_cost"] == both[q]["total_cost"]
    # elite costs ascend; the best path is the cheapest of the nSave back-traced paths (trajectory.py:125-148), which
    # need not end in the cheapest elite of the last window
    assert (np.diff(both[0]["elite_cost"], axis=1) >= 0).all()
    assert both[0]["total_cost"] == both[0]["path_total_costs"].min()
    # Q = 1 with host noise is the reference loop: same best path as SamCon / Trajectory.find_paths on the same pool
    # arithmetic (the batched pool stores float32 rows, hence the tolerance), over enough windows for paths to cross
    kw = dict(nIter=4, nSample=12, nSave=3)
    one = BatchedSamCon(cfg, [loaders[0]], pool, device_noise=False, **kw).sample()[0]
    cfg.motion_seq = "x"
    ref = SamCon(cfg, loaders[0], 1, 4, 12, 3, pool=OraclePool(oracle))
    out, info = ref.sample(save=False)
    np.testing.assert_allclose(out["x"]["path_0"]["cost"].reshape(-1), one["cost"], rtol=1e-4)
    np.testing.assert_allclose(out["x"]["path_0"]["total_cost"], one["total_cost"], rtol=1e-4)
    np.testing.assert_allclose(out["x"]["path_0"]["action"], one["action"], atol=1e-6)
    np.testing.assert_allclose(sorted(out["x"][f"path_{i}"]["total_cost"] for i in range(3)), np.sort(one["path_total_costs"]), rtol=1e-4)


def test_result_pickles_have_the_reference_layout(oracle, motion, tmp_path):
    """find_path_and_save writes the two joblib files the reference writes (trajectory.py:151-201): same directory
    layout, keys, shapes and dtypes, so scripts/eval_samcon.py can consume them unchanged."""
    import glob
    import joblib
    (name, _), = motion.items()
    cfg = Config("walk", base_dir=str(tmp_path), cfg_dict=Config("walk").cfg_dict)
    cfg.motion_seq = name
    sc = SamCon(cfg, AmassMotionData(motion, name), 1, nIter=2, nSample=12, nSave=3, pool=OraclePool(oracle))
    sc.sample(save=True)
    (res_file,) = glob.glob(str(tmp_path / "results" / "samcon" / "walk" / "results" / "walk_*.pkl"))
    (info_file,) = glob.glob(str(tmp_path / "results" / "samcon" / "walk" / "info" / "walk_*.pkl"))
    out, info = joblib.load(res_file), joblib.load(info_file)
    assert list(out) == [name] and sorted(out[name]) == ["path_0", "path_1", "path_2"]
    p = out[name]["path_0"]
    assert sorted(p) == ["action", "cost", "simulated_state", "t0_state", "target_state", "total_cost"]
    assert p["target_state"].shape == (2, 132) and p["simulated_state"].shape == (2, 132)
    assert p["action"].shape == (2, 68) and p["cost"].shape == (2, 1) and p["t0_state"].shape == (132,)
    assert np.isclose(p["total_cost"], p["cost"].sum()) and all(v.dtype == np.float64 for k, v in p.items() if k != "total_cost")
    costs = [out[name][f"path_{i}"]["total_cost"] for i in range(3)]
    assert costs == sorted(costs)                                          # paths ordered by total cost (trajectory.py:147)
    assert sorted(info) == ["balance_cost", "best_path_my_sys", "com_cost", "ee_cost", "elite_inds", "pose_cost",
                            "root_cost", "total_cost"]
    for k in ("pose_cost", "root_cost", "ee_cost", "balance_cost", "com_cost", "total_cost"):
        assert info[k].shape == (2, 12)
    assert info["elite_inds"].shape == (2, 3) and info["best_path_my_sys"].shape == (2,)

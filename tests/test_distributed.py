"""The sharded search (samcon_b200/distributed.py) on 2 ranks over gloo, CPU only: rank r rolls out its block of the
window's samples (oracle-backed pool stand-in), local top-E -> all_gather of (cost, index) -> identical final
select on every rank -> winners' states exchanged.  The 2-rank result must equal the unsharded search."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    sys.path.insert(0, HERE)
    from helpers import near_action, stand
    rng = np.random.default_rng(21)
    E, children = 4, 3
    parents = np.array([stand(rng, 0.05, 0.2, 0.99) for _ in range(E)])
    actions = np.array([near_action(parents[k // children], rng, 0.2) for k in range(E * children)])
    target = stand(rng)
    return parents, actions, target, E, children


def _make_pool():
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle.oracle import Oracle
    from oracle_pool import OracleDevicePool
    from samcon_b200.config import Config
    from samcon_b200.model import Character
    cfg = Config("walk")
    return OracleDevicePool(Oracle(Character.from_config(cfg)), cfg)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from samcon_b200.distributed import ShardedSearch
        pool = _make_pool()
        parents, actions, target, E, children = _inputs()
        S = len(actions)
        S_loc, E_loc = S // world, E // world
        search = ShardedSearch(pool, S_loc, E, rank, world)
        assert search.exchange == "all_to_all"                   # no symmetric memory on CPU tensors: the collective path
        f = lambda x: torch.tensor(x, dtype=torch.float32)
        # elites are dealt round robin; local sample j is sample gid[j] of the unsharded search
        gid = search.global_sample_ids(rank, torch.arange(S_loc)).numpy()
        acts = f(actions[gid])
        blk, bcache, ids, ecost, _ = search.window(f(parents[rank::world]), acts, f(target), 15)
        # second window from the exchanged parent block, contact caches included
        blk2, bcache2, ids2, ecost2, _ = search.window(blk, acts, f(target), 15, parent_cache=bcache)
        out[rank] = (blk.numpy(), bcache.numpy(), ids.numpy(), ecost.numpy(), blk2.numpy(), bcache2.numpy(), ids2.numpy(), ecost2.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_search_equals_unsharded():
    """Two chained windows on 2 ranks == the same two windows on one: elite ids, costs, parent states and contact caches."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    pool = _make_pool()
    parents, actions, target, E, children = _inputs()
    f = lambda x: torch.tensor(x, dtype=torch.float32)
    E_loc = E // world
    st, cost, info, cache = pool.window(f(parents), f(actions), f(target), 15, out_cache=True)
    idx, ec = pool.select(cost, E)
    ref_ids, ref_states, ref_cache = idx[0].numpy(), st[idx[0].long()], cache[idx[0].long()]
    st2, cost2, _, cache2 = pool.window(ref_states, f(actions), f(target), 15, parent_cache=ref_cache, out_cache=True)
    idx2, ec2 = pool.select(cost2, E)
    assert (ref_cache[:, 0] >= 0).any()                          # the scenario does carry contacts across the windows
    for r in range(world):
        blk, bcache, ids, ecost, blk2, bcache2, ids2, ecost2 = out[r]
        assert list(ids) == list(ref_ids)                        # global sample ids of the elites, ascending cost
        np.testing.assert_array_equal(ecost, ec[0].numpy())
        np.testing.assert_array_equal(blk, ref_states[r::world].numpy())                       # x + 0 exchange is exact
        np.testing.assert_array_equal(bcache, ref_cache[r::world].numpy())
        assert list(ids2) == list(idx2[0].numpy())
        np.testing.assert_array_equal(ecost2, ec2[0].numpy())
        np.testing.assert_array_equal(blk2, st2[idx2[0].long()][r::world].numpy())
        np.testing.assert_array_equal(bcache2, cache2[idx2[0].long()][r::world].numpy())


def _seq_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from samcon_b200.config import Config
        from samcon_b200.distributed import solve_sequences_sharded
        from samcon_b200.motion import AmassMotionData, make_synthetic_motion
        loaders = []
        for seed in (0, 1, 2):
            m = make_synthetic_motion("gait", T=12, seed=seed)
            (name, _), = m.items()
            loaders.append(AmassMotionData(m, name))
        q0, res = solve_sequences_sharded(Config("walk"), loaders, _make_pool(), rank, world, nIter=2, nSample=6, nSave=2)
        out[rank] = (q0, [(r["total_cost"], r["action"]) for r in res])
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sequences_sharded_over_ranks():
    """Config 5 on N ranks: sequences are dealt to the ranks in blocks, no collective; every sequence gets the result it
    gets when all are solved by one rank (its Philox stream is keyed by its global index)."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_seq_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    one = mgr.dict()
    mp.spawn(_seq_worker, args=(1, _free_port(), one), nprocs=1, join=True)
    assert out[0][0] == 0 and out[1][0] == 2 and len(out[0][1]) == 2 and len(out[1][1]) == 1
    both = out[0][1] + out[1][1]
    for (c, a), (c1, a1) in zip(both, one[0][1]):
        assert c == c1
        np.testing.assert_array_equal(a, a1)


def _solve_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from samcon_b200.config import Config
        from samcon_b200.distributed import ShardedSamCon
        from samcon_b200.motion import AmassMotionData, make_synthetic_motion
        m = make_synthetic_motion("gait", T=12, seed=0)
        (name, _), = m.items()
        res = ShardedSamCon(Config("walk"), AmassMotionData(m, name), _make_pool(), rank, world, nIter=3, nSample=12, nSave=4).sample()
        out[rank] = None if res is None else {k: res[k] for k in ("simulated_state", "action", "cost", "total_cost", "elite_cost", "path_total_costs")}
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_solve_best_path():
    """ShardedSamCon on 2 ranks: rank 0 gets a best path that is a chain (each window's state is the result of the stored
    action from the previous window's state, contact cache included), its total is the cheapest of the back-traced totals,
    and the other rank returns nothing."""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_solve_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[1] is None
    r = out[0]
    assert r["simulated_state"].shape == (3, 132) and r["action"].shape == (3, 68) and r["elite_cost"].shape == (3, 4)
    assert (np.diff(r["elite_cost"], axis=1) >= 0).all()
    assert r["total_cost"] == r["path_total_costs"].min() and np.isclose(r["cost"].sum(), r["total_cost"])
    # replay the path on the oracle: window by window from the path's own previous state (float32 rows, hence the tolerance)
    from samcon_b200.config import Config
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    sys.path.insert(0, HERE)
    from helpers import joint_rms
    pool = _make_pool()
    m = make_synthetic_motion("gait", T=12, seed=0)
    (name, _), = m.items()
    ld = AmassMotionData(m, name)
    s, cache = ld.getSpecTimeState(0.0), None
    for t in range(3):
        st, c, _, cc = pool.o.window(np.asarray(s, np.float32).astype(np.float64)[None], r["action"][t][None], np.asarray(ld.getSpecTimeState(0.1 * (t + 1)), np.float32).astype(np.float64), 100,
                                     parent_cache=cache, return_cache=True)
        assert joint_rms(st, r["simulated_state"][t][None])[0] < 1e-5
        assert abs(c[0] - r["cost"][t]) < 1e-4 * max(1.0, c[0])
        s, cache = r["simulated_state"][t], cc

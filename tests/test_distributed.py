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
        f = lambda x: torch.tensor(x, dtype=torch.float32)
        blk, ids, ecost, (st, cost, info) = search.window(f(parents[rank * E_loc:(rank + 1) * E_loc]),
                                                          f(actions[rank * S_loc:(rank + 1) * S_loc]), f(target), 15)
        # second window from the exchanged elites: every rank must hold the same elite table
        elites, ids2, _ = search.select_and_exchange(st, cost)
        out[rank] = (blk.numpy(), ids.numpy(), ecost.numpy(), elites.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_search_equals_unsharded():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    pool = _make_pool()
    parents, actions, target, E, children = _inputs()
    f = lambda x: torch.tensor(x, dtype=torch.float32)
    st, cost, info = pool.window(f(parents), f(actions), f(target), 15)
    idx, ec = pool.select(cost, E)
    ref_ids = idx[0].numpy()
    ref_states = st[idx[0].long()].numpy()
    for r in range(world):
        blk, ids, ecost, elites = out[r]
        assert list(ids) == list(ref_ids)                        # global sample ids of the elites, ascending cost
        np.testing.assert_array_equal(ecost, ec[0].numpy())
        np.testing.assert_array_equal(elites, ref_states)        # x + 0 exchange is exact
        np.testing.assert_array_equal(blk, ref_states[r * (E // world):(r + 1) * (E // world)])
    np.testing.assert_array_equal(out[0][3], out[1][3])

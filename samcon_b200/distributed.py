"""Samples of one SamCon window sharded over the GPUs of one box (one process per GPU).

The reference shards the nSave elites contiguously over ``num_processes`` workers and gathers pickled result
lists over pipes (/root/reference/samcon/core/samcon.py:85-90, 168-176).  Here the elites -- which come out of the
selection in ascending cost order -- are dealt to the ranks ROUND ROBIN: rank r simulates the children of elites
r, r + N, r + 2N, ...  (Contiguous blocks, round 1's choice, hand rank 0 the cheapest elites and rank N-1 the costliest,
and costly parents are the ones with many contacts: on 4 GPUs the last rank's rollouts were 0.9 ms per window slower
than the first's and every window waits for the slowest rank.)  Local sample j of rank r is sample
((j // children) * N + r) * children + j % children of the unsharded search (sample k descends from elite k // children,
samcon.py:31-35), so a sharded search fed the same actions is the unsharded one.  Rollouts need no communication.  The
one exchange step per window is elite selection:

  1. ONE NCCL all_gather of the ranks' cost vectors (4 B per sample: 40 KB per rank at 10 k samples; the index of a cost
     is its position, so (cost, index) pairs need not travel),
  2. every rank runs the identical top-E over the N*S_loc costs with the cooperative radix select (ties: lower rank, then
     lower local index) -- one selection instead of round 1's local top-E, all_gather of (cost, index) lists and final
     top-E over N*E candidates, which cost 0.11 + 0.1 + 0.14 ms per window on 4 GPUs,
  3. owner rank and row of every winner follow from its position,
  4. rank r fetches the snapshots (132-float state + contact cache) of ITS parents -- E_loc rows, wherever they
     were simulated -- straight out of the owners' result buffers over NVLink: the buffers live in symmetric memory
     (torch.distributed._symmetric_memory, CUDA VMM handles exchanged once at start-up) and one gather kernel
     (sc_gather_rows_peer) reads peer HBM through NVSwitch.  No collective, no padding, no host synchronisation:
     E_loc x 592 B per rank per window.  The all_gather of step 2 is what orders the traffic: a rank's contribution
     is sent after its rollout finished (same stream), so once the all_gather has completed every peer's buffer is
     complete; result buffers alternate between two copies from window to window so that a fast rank's next rollout
     never overwrites rows a slow rank is still reading (the next all_gather cannot complete before the slow rank
     has enqueued its contribution, which follows its reads in stream order).
     Where symmetric memory is unavailable (gloo in the CPU tests, or a torch without it) the rows travel by ONE
     all_to_all_single of fixed-size blocks instead (each rank sends every peer the rows of that peer's block it owns,
     zero elsewhere; the receiver adds the N blocks -- every row has exactly one owner, x + 0 is exact).

All steps are enqueued on the current stream; there is no host synchronisation inside a window.
``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .abi import SC_CACHE_DIM as CACHE_DIM
from .model import STATE_DIM

ROW = STATE_DIM + CACHE_DIM      # a sample's snapshot: state + contact cache


class ShardedSearch:
    def __init__(self, pool, S_local: int, E_global: int, rank: int, world: int, group=None, exchange: str = "auto"):
        if E_global % world or S_local % (E_global // world):
            raise ValueError("nSave must split evenly over ranks and nSample over elites")
        self.pool, self.S_loc, self.E, self.rank, self.world, self.group = pool, S_local, E_global, rank, world, group
        self.E_loc = E_global // world
        self.children = S_local // self.E_loc
        self.exchange = "local" if world == 1 else None
        self._win = 0
        self._snap = None
        dev = getattr(pool, "device", torch.device("cpu"))
        if world > 1 and exchange in ("auto", "p2p") and dev.type == "cuda" and hasattr(pool, "gather_rows_peer"):
            try:
                self._setup_symmetric(dev)
                self.exchange = "nvlink-p2p"
            except Exception as e:                # noqa: BLE001 -- any failure means: use the collective
                if exchange == "p2p":
                    raise
                self._symm_error = repr(e)
        if self.exchange is None:
            self.exchange = "all_to_all"
        if self._snap is None:
            n = self.S_loc * ROW
            self._snap = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]

    # ------------------------------------------------------------------ symmetric result buffers
    def _setup_symmetric(self, dev):
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        n = self.S_loc * ROW
        self._snap, self._peer_state, self._peer_cache, self._hdl = [], [], [], []
        for _ in range(2):
            t = symm.empty(n, dtype=torch.float32, device=dev)
            hdl = symm.rendezvous(t, group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != self.world:
                raise RuntimeError("symmetric memory rendezvous returned the wrong number of peers")
            self._snap.append(t)
            self._hdl.append(hdl)
            self._peer_state.append(torch.tensor(ptrs, dtype=torch.int64, device=dev))
            self._peer_cache.append(torch.tensor([p + 4 * self.S_loc * STATE_DIM for p in ptrs], dtype=torch.int64, device=dev))

    def result_buffers(self):
        """(state [S_loc,132], cache [S_loc,CACHE_DIM]) views of this window's snapshot buffer."""
        b = self._snap[self._win & 1]
        ns = self.S_loc * STATE_DIM
        return b[:ns].view(self.S_loc, STATE_DIM), b[ns:].view(self.S_loc, CACHE_DIM)

    def global_sample_ids(self, rank, local):
        """Index in the unsharded search of local sample `local` of rank `rank` (tensors or ints)."""
        return ((local // self.children) * self.world + rank) * self.children + local % self.children

    def my_slots(self):
        """Positions in the global (ascending-cost) elite list of the parents this rank simulates."""
        return slice(self.rank, self.E, self.world)

    # ------------------------------------------------------------------ the exchange
    def select_global(self, cost):
        """all_gather of the cost vectors, identical top-E on every rank.
        -> (owner rank [E] int32, row in the owner's buffers [E] int32, cost [E], global sample id [E] int64)."""
        dev = cost.device
        if self.world == 1:
            idx, ec = self.pool.select(cost, self.E)
            idx = idx.reshape(-1)
            return torch.zeros(self.E, dtype=torch.int32, device=dev), idx.contiguous(), ec.reshape(-1), idx.to(torch.int64)
        all_cost = torch.empty((self.world * self.S_loc,), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(all_cost, cost.contiguous(), group=self.group)
        pos, gcost = self.pool.select(all_cost, self.E)
        pos = pos.reshape(-1)
        win_rank = torch.div(pos, self.S_loc, rounding_mode="floor").to(torch.int32)
        win_local = (pos - win_rank * self.S_loc).to(torch.int32).contiguous()
        return win_rank, win_local, gcost.reshape(-1), self.global_sample_ids(win_rank.to(torch.int64), win_local.to(torch.int64))

    def fetch_block(self, win_rank, win_local, st, cache):
        """Snapshots of this rank's parents (elites rank, rank + N, ...): (states [E_loc,132], caches [E_loc,CACHE_DIM])."""
        sl = self.my_slots()
        if self.world == 1:
            return self.pool.gather_rows(st, win_local[sl]), self.pool.gather_rows(cache, win_local[sl])
        if self.exchange == "nvlink-p2p":
            p = self._win & 1
            r, l = win_rank[sl].contiguous(), win_local[sl].contiguous()
            return (self.pool.gather_rows_peer(self._peer_state[p], r, l, STATE_DIM),
                    self.pool.gather_rows_peer(self._peer_cache[p], r, l, CACHE_DIM))
        # collective fall-back: block d of the send buffer = the rows of rank d's parents that live here
        dev = st.device
        mine = win_rank == self.rank
        safe = torch.where(mine, win_local, torch.zeros_like(win_local)).to(torch.int32)
        rows = torch.cat([self.pool.gather_rows(st, safe), self.pool.gather_rows(cache, safe)], dim=1)      # [E, ROW]
        rows = torch.where(mine[:, None], rows, torch.zeros((), dtype=torch.float32, device=dev))
        send = rows.view(self.E_loc, self.world, ROW).transpose(0, 1).contiguous()         # [d, j] = elite j * N + d
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1, ROW), send.view(-1, ROW), group=self.group)    # recv[s, j] = rank s's copy of my parent j
        blk = recv.sum(dim=0)
        return blk[:, :STATE_DIM].contiguous(), blk[:, STATE_DIM:].contiguous()

    def select_and_exchange(self, st, cost, cache=None):
        """(st [S_loc,132], cost [S_loc], cache [S_loc,CACHE_DIM]) of this rank -> (this rank's next parents [E_loc,132],
        their caches, elite sample ids in the unsharded search's numbering [E], elite costs [E]); ids and costs identical
        on every rank."""
        if cache is None:
            cache = torch.zeros((st.shape[0], CACHE_DIM), dtype=torch.float32, device=st.device)
            cache[:, :CACHE_DIM // 2] = -1.0
        win_rank, win_local, gcost, gid = self.select_global(cost)
        bst, bcache = self.fetch_block(win_rank, win_local, st, cache)
        self._win += 1
        return bst, bcache, gid, gcost

    def window(self, parent_block, actions, target, nsteps, parent_cache=None):
        """One sharded window.  parent_block [E_loc,132] (+ parent_cache [E_loc,CACHE_DIM]) are this rank's parents.
        Returns (next parent block, next parent caches, elite ids [E], elite costs [E], (st, cost, info) of the local
        samples -- st is a view of this window's snapshot buffer, valid until the window after next)."""
        st, cache = self.result_buffers()
        dev = st.device
        cost = torch.empty((self.S_loc,), dtype=torch.float32, device=dev)
        info = torch.empty((self.S_loc, 5), dtype=torch.float32, device=dev)
        self.pool.window(parent_block, actions, target, nsteps, out=(st, cost, info), parent_cache=parent_cache, out_cache=cache)
        bst, bcache, ids, ecost = self.select_and_exchange(st, cost, cache)
        return bst, bcache, ids, ecost, (st, cost, info)


def solve_sequences_sharded(cfg, data_loaders, pool, rank: int, world: int, **kw):
    """BASELINE config 5 on N GPUs: the Q sequences are dealt to the ranks in contiguous blocks and every rank solves its
    share with BatchedSamCon -- sequences are independent searches, so there is no data-path collective at all.
    Sequence q draws from the Philox stream (cfg.seed + q) whichever rank solves it.  -> (first sequence index, results)."""
    from .batched import BatchedSamCon
    Q = len(data_loaders)
    per = (Q + world - 1) // world
    q0, q1 = min(Q, rank * per), min(Q, (rank + 1) * per)
    if q1 <= q0:
        return q0, []
    return q0, BatchedSamCon(cfg, data_loaders[q0:q1], pool, seed_offset=q0, **kw).sample()


class ShardedSamCon:
    """A whole SamCon solve of ONE sequence with the window's samples sharded over the ranks (BASELINE config 3: "solve
    sharded over 8 GPUs"): the reference's loop (/root/reference/samcon/core/samcon.py:144-198) with ``ShardedSearch.window``
    in the place of the worker pool.  Every rank draws the perturbations of its own samples on the device (Philox stream
    cfg.seed + 1000 * rank, window t); per window rank 0 additionally receives the E elites' (state, action) rows -- one
    reduce of a zero-padded [E, 200] buffer, outside the selection's critical path -- and keeps the history the best-path
    back-trace needs.  ``sample()`` returns, on rank 0, the per-path arrays of the reference's result dict for the best path
    (``simulated_state``, ``action``, ``cost``, ``target_state``, ``t0_state``, ``total_cost``) plus ``elite_cost``; None on the
    other ranks."""

    def __init__(self, cfg, data_loader, pool, rank, world, nIter=None, nSample=None, nSave=None, group=None):
        self.cfg, self.loader, self.pool, self.rank, self.world, self.group = cfg, data_loader, pool, rank, world, group
        self.nIter = cfg.nIter if nIter is None else nIter
        self.S = cfg.nSample if nSample is None else nSample          # nSample of the whole search
        self.E = cfg.nSave if nSave is None else nSave
        if self.S % self.E != 0:
            raise Exception("nSample should be a multiples of nSave")          # samcon.py:66-67
        if self.S % world or self.E % world:
            raise ValueError("nSample and nSave must split evenly over the ranks")
        if cfg.fps_sim % cfg.fps_act != 0:
            raise Exception("FPS_SIM should be a multiples of FPS_ACT")
        self.steps = cfg.fps_sim // cfg.fps_act
        self.search = ShardedSearch(pool, self.S // world, self.E, rank, world, group)

    def sample(self):
        import numpy as np
        pool, search, dev = self.pool, self.search, getattr(self.pool, "device", torch.device("cpu"))
        dt = 1.0 / self.cfg.fps_act
        targets = np.stack([self.loader.getSpecTimeState(t * dt) for t in range(self.nIter + 1)])
        d_targets = torch.tensor(targets, dtype=torch.float32, device=dev)
        limits = torch.tensor(np.asarray(self.cfg.values), dtype=torch.float32, device=dev)
        parents = d_targets[0:1].repeat(search.E_loc, 1).contiguous()
        pcache = None
        children = self.S // self.E
        hist = []
        for t in range(1, self.nIter + 1):
            act = pool.sample_gen(d_targets[t], limits, search.S_loc, gaussian=self.cfg.noise_type == "gaussian",
                                  seed=self.cfg.seed + 1000 * self.rank, window=t)
            st, cache = search.result_buffers()
            cost = torch.empty((search.S_loc,), dtype=torch.float32, device=dev)
            info = torch.empty((search.S_loc, 5), dtype=torch.float32, device=dev)
            pool.window(parents, act, d_targets[t], self.steps, out=(st, cost, info), parent_cache=pcache, out_cache=cache)
            win_rank, win_local, gcost, gid = search.select_global(cost)
            parents, pcache = search.fetch_block(win_rank, win_local, st, cache)
            search._win += 1
            # the elites' rows for the record: every rank fills in the ones it simulated, rank 0 receives the sum
            mine = win_rank == self.rank
            safe = torch.where(mine, win_local, torch.zeros_like(win_local)).to(torch.int32)
            rows = torch.cat([pool.gather_rows(st, safe), pool.gather_rows(act, safe)], dim=1)
            rows = torch.where(mine[:, None], rows, torch.zeros((), dtype=torch.float32, device=dev)).contiguous()
            if self.world > 1:
                dist.reduce(rows, dst=0, op=dist.ReduceOp.SUM, group=self.group)
            if self.rank == 0:
                hist.append((rows[:, :STATE_DIM].cpu().numpy().astype(np.float64), rows[:, STATE_DIM:].cpu().numpy().astype(np.float64),
                             gcost.cpu().numpy().astype(np.float64), (gid // children).cpu().numpy()))
        if self.rank != 0:
            return None
        # best path: back-trace every elite of the last window through the parent slots, cheapest total (trajectory.py:118-154)
        n, E = len(hist), self.E
        slots = np.zeros((n, E), dtype=np.int64)
        slots[-1] = np.arange(E)
        for t in range(n - 1, 0, -1):
            slots[t - 1] = hist[t][3][slots[t]]
        total = sum(hist[t][2][slots[t]] for t in range(n))
        j = int(np.argsort(total, kind="stable")[0])
        pick = lambda k: np.array([hist[t][k][slots[t, j]] for t in range(n)])
        return {"simulated_state": pick(0), "action": pick(1), "cost": pick(2), "target_state": targets[1:], "t0_state": targets[0],
                "total_cost": float(total[j]), "elite_cost": np.stack([h[2] for h in hist]), "path_total_costs": total}

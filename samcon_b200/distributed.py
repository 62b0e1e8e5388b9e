"""Samples of one SamCon window sharded over the GPUs of one box (one process per GPU).

The reference shards the nSave elites contiguously over ``num_processes`` workers and gathers pickled result
lists over pipes (/root/reference/samcon/core/samcon.py:85-90, 168-176).  Here rank r owns samples
[r*S_loc, (r+1)*S_loc) and therefore -- sample k descends from elite k // children (samcon.py:31-35) -- the
contiguous elite block [r*E_loc, (r+1)*E_loc).  Rollouts need no communication.  The one exchange step per
window is elite selection:

  1. every rank takes its local top-min(E, S_loc) (cost, index) with the segmented radix top-k kernel,
  2. NCCL all_gather of those (cost f32, index i32) lists  (8 B x E per rank),
  3. every rank runs the identical final top-E over the N*E candidates (ties: lower global sample index),
  4. winners' states are exchanged: each rank writes the rows it owns into a zero [E,132] buffer and one
     NCCL all_reduce(sum) delivers every winning state to every rank (x + 0 is exact), of which rank r keeps
     its parent block.

All steps are enqueued on the current stream; there is no host synchronisation inside a window.
``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .model import STATE_DIM


class ShardedSearch:
    def __init__(self, pool, S_local: int, E_global: int, rank: int, world: int, group=None):
        if E_global % world or S_local % (E_global // world):
            raise ValueError("nSave must split evenly over ranks and nSample over elites")
        self.pool, self.S_loc, self.E, self.rank, self.world, self.group = pool, S_local, E_global, rank, world, group
        self.E_loc = E_global // world
        self.Ec = min(E_global, S_local)          # candidates a rank can contribute

    def select_and_exchange(self, st, cost):
        """(st [S_loc,132], cost [S_loc]) of this rank -> (all elite states [E,132], elite global sample ids [E],
        elite costs [E]); identical on every rank."""
        dev = cost.device
        lidx, lcost = self.pool.select(cost, self.Ec)
        lidx, lcost = lidx.reshape(-1), lcost.reshape(-1)
        if self.world == 1:
            return self.pool.gather_rows(st, lidx[:self.E]), lidx[:self.E].to(torch.int64), lcost[:self.E]
        # flat [world * Ec] outputs: the layout both NCCL and gloo accept for a 1-D input
        all_cost = torch.empty((self.world * self.Ec,), dtype=torch.float32, device=dev)
        all_idx = torch.empty((self.world * self.Ec,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_cost, lcost.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_idx, lidx.contiguous(), group=self.group)
        pos, gcost = self.pool.select(all_cost, self.E)
        pos = pos.reshape(-1).to(torch.int64)
        win_rank = pos // self.Ec
        win_local = all_idx.reshape(-1)[pos].to(torch.int64)
        mine = win_rank == self.rank
        buf = torch.zeros((self.E, STATE_DIM), dtype=torch.float32, device=dev)
        # rows this rank owns; the masked local index is clamped so the gather stays in range
        rows = self.pool.gather_rows(st, torch.where(mine, win_local, torch.zeros_like(win_local)).to(torch.int32))
        buf = torch.where(mine[:, None], rows, buf)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf, win_rank * self.S_loc + win_local, gcost.reshape(-1)

    def window(self, parent_block, actions, target, nsteps):
        """One sharded window.  parent_block [E_loc,132] are this rank's parents.  Returns
        (next parent block [E_loc,132], elite ids [E], elite costs [E], (st, cost, info) of the local samples)."""
        st, cost, info = self.pool.window(parent_block, actions, target, nsteps)
        elites, ids, ecost = self.select_and_exchange(st, cost)
        blk = elites[self.rank * self.E_loc:(self.rank + 1) * self.E_loc].contiguous()
        return blk, ids, ecost, (st, cost, info)

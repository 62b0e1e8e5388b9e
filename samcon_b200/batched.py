"""Many SamCon searches at once: Q motion sequences solved concurrently, device resident.

The reference solves one sequence per ``run_samcon.py`` invocation (/root/reference/scripts/run_samcon.py:19-40,
one worker pool per run).  The windows of different sequences are independent, so here the samples of all Q
sequences of window t go through ONE ``sc_window`` launch (sample k belongs to sequence k // nSample, is scored
against that sequence's target and descends from that sequence's elite (k % nSample) // children), ONE segmented
``sc_select`` keeps every sequence's nSave best, and ONE ``sc_gather_rows`` makes them the next parents.  Nothing
returns to the host inside the loop: per window only the elites' (state, action, cost, parent slot) are kept (on
the device), which is all the reference's best-path extraction needs
(/root/reference/samcon/core/trajectory.py:118-154: walk the uid tree back from the elites of the last window).

With Q = 1 and ``device_noise=False`` this is the reference loop (same MT19937 draws as ``SamCon``); with
``device_noise=True`` the perturbations are Philox draws made on the device (``sc_sample_gen``).
"""
from __future__ import annotations

import numpy as np
import torch

from .model import STATE_DIM
from .sampling import get_batch_action


class BatchedSamCon:

    def __init__(self, cfg, data_loaders, pool, nIter=None, nSample=None, nSave=None, device_noise=True, search=None):
        """``data_loaders``: one AmassMotionData-like object per sequence (``getSpecTimeState(t)``);
        ``pool``: a GpuSamplePool;  ``search``: optional ``distributed.ShardedSearch`` factory is not needed here --
        sequences, not samples, are what one would shard across GPUs for this mode."""
        self._cfg = cfg
        self._loaders = list(data_loaders)
        self._pool = pool
        self.Q = len(self._loaders)
        self.nIter = cfg.nIter if nIter is None else nIter
        self.S = cfg.nSample if nSample is None else nSample
        self.E = cfg.nSave if nSave is None else nSave
        if self.S % self.E != 0:
            raise Exception("nSample should be a multiples of nSave")          # samcon.py:66-67
        if cfg.fps_sim % cfg.fps_act != 0:
            raise Exception("FPS_SIM should be a multiples of FPS_ACT")        # humanoid_tracking_env.py:27-28
        self.children = self.S // self.E
        self.steps = cfg.fps_sim // cfg.fps_act
        self._device_noise = device_noise
        self._rngs = [np.random.RandomState(cfg.seed) for _ in range(self.Q)]  # np.random.seed(cfg.seed) per run
        dev = pool.device
        Q, S, E = self.Q, self.S, self.E
        k = torch.arange(Q * S, device=dev, dtype=torch.int32)
        self._sample_seq = (k // S).contiguous()
        self._parent_idx = ((k // S) * E + (k % S) // self.children).contiguous()
        self._seg = (torch.arange(Q + 1, device=dev, dtype=torch.int32) * S).contiguous()
        dt = 1.0 / cfg.fps_act
        tg = np.stack([np.stack([ld.getSpecTimeState(t * dt) for t in range(self.nIter + 1)]) for ld in self._loaders])
        self.targets_host = tg                                               # [Q, nIter+1, 132]
        self._targets = torch.tensor(tg, dtype=torch.float32, device=dev)
        self._limits = torch.tensor(np.asarray(cfg.values), dtype=torch.float32, device=dev)

    def _actions(self, t):
        Q, S = self.Q, self.S
        if self._device_noise:
            act = torch.empty((Q * S, self._pool.act_dim), dtype=torch.float32, device=self._pool.device)
            for q in range(Q):
                self._pool.sample_gen(self._targets[q, t], self._limits, S, gaussian=self._cfg.noise_type == "gaussian",
                                      seed=self._cfg.seed + q, window=t, out=act[q * S:(q + 1) * S])
            return act
        host = np.concatenate([get_batch_action(self.targets_host[q, t], S, self._cfg.values, self._cfg.noise_type,
                                                self._rngs[q]) for q in range(Q)])
        return torch.tensor(host, dtype=torch.float32, device=self._pool.device)

    def sample(self):
        """Run all windows.  Returns a list (one per sequence) of dicts with the reference's per-path arrays for the
        best path: ``simulated_state`` (nIter,132), ``action`` (nIter,68), ``cost`` (nIter,), ``target_state``
        (nIter,132), ``t0_state`` (132,), ``total_cost``; plus ``elite_cost`` (nIter, nSave)."""
        pool, Q, S, E = self._pool, self.Q, self.S, self.E
        parents = self._targets[:, 0, :].repeat_interleave(E, dim=0).contiguous()          # every elite slot = t0 state
        hist = []
        for t in range(1, self.nIter + 1):
            act = self._actions(t)
            st, cost, info = pool.window(parents, act, self._targets[:, t, :].contiguous(), self.steps,
                                         parent_idx=self._parent_idx, sample_seq=self._sample_seq)
            idx, ecost = pool.select(cost, E, seg_offsets=self._seg)                       # [Q,E] global sample ids
            flat = idx.reshape(-1)
            parents = pool.gather_rows(st, flat)
            eact = pool.gather_rows(act, flat)
            eparent = (flat.to(torch.int64) % S) // self.children                         # elite slot of window t-1
            hist.append((parents, eact, ecost, eparent.reshape(Q, E)))
        # ---- best path per sequence: back-trace from the cheapest elite of the last window (trajectory.py:118-154)
        H = [(p.reshape(Q, E, STATE_DIM).cpu().numpy().astype(np.float64), a.reshape(Q, E, -1).cpu().numpy().astype(np.float64),
              c.cpu().numpy().astype(np.float64), e.cpu().numpy()) for p, a, c, e in hist]
        out = []
        for q in range(Q):
            e = 0
            sim, acts, costs = [], [], []
            for t in range(self.nIter - 1, -1, -1):
                P, A, Cc, Ep = H[t]
                sim.append(P[q, e]); acts.append(A[q, e]); costs.append(Cc[q, e])
                e = int(Ep[q, e])
            sim.reverse(); acts.reverse(); costs.reverse()
            out.append({"simulated_state": np.array(sim), "action": np.array(acts), "cost": np.array(costs),
                        "target_state": self.targets_host[q, 1:], "t0_state": self.targets_host[q, 0],
                        "total_cost": float(np.sum(costs)), "elite_cost": np.stack([h[2][q] for h in H])})
        return out

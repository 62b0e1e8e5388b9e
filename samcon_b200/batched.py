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

    def __init__(self, cfg, data_loaders, pool, nIter=None, nSample=None, nSave=None, device_noise=True, seed_offset=0):
        """``data_loaders``: one AmassMotionData-like object per sequence (``getSpecTimeState(t)``);
        ``pool``: a GpuSamplePool.  Sequences, not samples, are what is sharded across GPUs in this mode
        (``distributed.solve_sequences_sharded``): every rank runs this class on its share, no collective."""
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
        self._seed_offset = int(seed_offset)      # first sequence's stream: a rank that solves sequences [q0, q0+Q) passes q0
        self._rngs = [np.random.RandomState(cfg.seed) for _ in range(self.Q)]  # np.random.seed(cfg.seed) per run
        dev = pool.device
        Q, S, E = self.Q, self.S, self.E
        k = torch.arange(Q * S, device=dev, dtype=torch.int32)
        self._sample_seq = (k // S).contiguous()
        self._parent_idx = ((k // S) * E + (k % S) // self.children).contiguous()
        self._seg = (torch.arange(Q + 1, device=dev, dtype=torch.int32) * S).contiguous()
        dt = 1.0 / cfg.fps_act
        times = np.arange(self.nIter + 1, dtype=np.float64) * dt
        if hasattr(pool, "motion_states") and all(hasattr(ld, "_motion") for ld in self._loaders):
            # slerp + finite differences for all Q x (nIter + 1) targets in one launch (sc_motion_states)
            self._targets = pool.motion_states(self._loaders, times)
            self.targets_host = self._targets.cpu().numpy().astype(np.float64)       # [Q, nIter+1, 132]
        else:
            tg = np.stack([np.stack([ld.getSpecTimeState(t) for t in times]) for ld in self._loaders])
            self.targets_host = tg
            self._targets = torch.tensor(tg, dtype=torch.float32, device=dev)
        self._limits = torch.tensor(np.asarray(cfg.values), dtype=torch.float32, device=dev)

    def _actions(self, t):
        Q, S = self.Q, self.S
        if self._device_noise:      # one launch for all sequences; sequence q draws from the stream (seed + q, t)
            return self._pool.sample_gen(self._targets[:, t, :].contiguous(), self._limits, S,
                                         gaussian=self._cfg.noise_type == "gaussian", seed=self._cfg.seed + self._seed_offset,
                                         window=t)
        host = np.concatenate([get_batch_action(self.targets_host[q, t], S, self._cfg.values, self._cfg.noise_type,
                                                self._rngs[q]) for q in range(Q)])
        return torch.tensor(host, dtype=torch.float32, device=self._pool.device)

    def begin(self):
        """Every elite slot of every sequence = the sequence's t0 state, empty contact caches, no history."""
        self._parents = self._targets[:, 0, :].repeat_interleave(self.E, dim=0).contiguous()
        self._pcache = None
        self._hist = []

    def step(self, t):
        """Window t (1-based) for all sequences: ONE sample_gen, ONE rollout, ONE segmented select, gathers.  Nothing
        returns to the host.  -> elite costs [Q,E] (device)."""
        pool, Q, S, E = self._pool, self.Q, self.S, self.E
        act = self._actions(t)
        st, cost, info, cache = pool.window(self._parents, act, self._targets[:, t, :].contiguous(), self.steps,
                                            parent_idx=self._parent_idx, sample_seq=self._sample_seq,
                                            parent_cache=self._pcache, out_cache=True)
        idx, ecost = pool.select(cost, E, seg_offsets=self._seg)                           # [Q,E] global sample ids
        flat = idx.reshape(-1)
        self._parents = pool.gather_rows(st, flat)
        self._pcache = pool.gather_rows(cache, flat)                                       # the snapshot carries the contact cache
        eact = pool.gather_rows(act, flat)
        eparent = (flat.to(torch.int64) % S) // self.children                             # elite slot of window t-1
        self._hist.append((self._parents, eact, ecost, eparent.reshape(Q, E)))
        return ecost

    def sample(self):
        """Run all windows.  Returns a list (one per sequence) of dicts with the reference's per-path arrays for the
        best path: ``simulated_state`` (nIter,132), ``action`` (nIter,68), ``cost`` (nIter,), ``target_state``
        (nIter,132), ``t0_state`` (132,), ``total_cost``; plus ``elite_cost`` (nIter, nSave)."""
        self.begin()
        for t in range(1, self.nIter + 1):
            self.step(t)
        return self.finish()

    def finish(self):
        Q, E = self.Q, self.E
        hist = self._hist
        nW = len(hist)                                             # windows stepped so far (nIter for a full solve)
        # ---- best path per sequence (trajectory.py:118-154): back-trace every elite of the last window through the
        # stored parent slots, sum the cost along each path, take the cheapest (stable ties, like find_paths' argsort)
        H = [(p.reshape(Q, E, STATE_DIM).cpu().numpy().astype(np.float64), a.reshape(Q, E, -1).cpu().numpy().astype(np.float64),
              c.cpu().numpy().astype(np.float64), e.cpu().numpy()) for p, a, c, e in hist]
        out = []
        for q in range(Q):
            slots = np.zeros((nW, E), dtype=np.int64)          # slots[t, j]: elite slot of path j in window t+1
            slots[-1] = np.arange(E)
            for t in range(nW - 1, 0, -1):
                slots[t - 1] = H[t][3][q, slots[t]]
            total = np.zeros(E)
            for t in range(nW):
                total += H[t][2][q, slots[t]]
            j = int(np.argsort(total, kind="stable")[0])
            sim = [H[t][0][q, slots[t, j]] for t in range(nW)]
            acts = [H[t][1][q, slots[t, j]] for t in range(nW)]
            costs = [H[t][2][q, slots[t, j]] for t in range(nW)]
            out.append({"simulated_state": np.array(sim), "action": np.array(acts), "cost": np.array(costs),
                        "target_state": self.targets_host[q, 1:nW + 1], "t0_state": self.targets_host[q, 0],
                        "total_cost": float(np.sum(costs)), "elite_cost": np.stack([h[2][q] for h in H]),
                        "path_total_costs": total})
        return out

"""GpuSamplePool -- the B200 replacement of the reference's master-worker pool.

The reference spawns ``num_processes`` workers, each owning a PyBullet world, and talks to them with
``["init"|"sim"|"close", data]`` messages (/root/reference/samcon/core/samcon.py:15-52, 70-90, 168-176);
sample states travel through ``.bullet`` files (/root/reference/envs/humanoid_tracking_env.py:178-182).
Here one object per GPU holds every sample's state in HBM and one ``sim`` call rolls *all* samples of a
window forward in a single kernel launch through the C ABI of include/samcon_b200.h.  PyTorch is used for
device memory and streams only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi
from .abi import SC_CACHE_DIM as CACHE_DIM
from .model import STATE_DIM, Character


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class GpuSamplePool:
    """All samples of one (or several) SamCon searches on one GPU.

    Low-level methods take and return CUDA tensors (float32 / int32, contiguous) and enqueue on the current
    torch stream without synchronising.  ``sim`` / ``init`` / ``close`` mirror the worker protocol and work on
    host numpy arrays like the reference's pipes do.
    """

    def __init__(self, character: Character, device: int | str | torch.device = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("GpuSamplePool needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.character = character
        self.nj = len(character.movable)
        self.act_dim = 4 * self.nj
        self._lib = abi.load()
        self._model, self._keep = abi.make_model(character)
        h = C.c_void_p()
        rc = self._lib.sc_create(C.byref(h), C.byref(self._model), self.device.index or 0)
        self._h = h
        if rc != 0:
            msg = self._lib.sc_last_error(h).decode() if h else "allocation failed"
            if h:
                self._lib.sc_destroy(h)
            self._h = None
            raise RuntimeError(f"sc_create failed ({rc}): {msg}")
        self._elites = None       # [E,132] device, parents of the next window
        self._elite_cache = None  # [E,CACHE_DIM] device, their contact caches (None: empty caches)
        self._closed = False

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self._lib.sc_last_error(self._h).decode()}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f32(self, x, shape=None):
        t = torch.as_tensor(x)
        if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _i32(self, x):
        if x is None:
            return None
        t = torch.as_tensor(x)
        if t.device != self.device or t.dtype != torch.int32 or not t.is_contiguous():
            t = t.to(device=self.device, dtype=torch.int32, non_blocking=True).contiguous()
        return t

    @property
    def launch_count(self):
        return int(self._lib.sc_launch_count(self._h))

    def rollout_config(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self._lib.sc_rollout_config(self._h, C.byref(a), C.byref(b), C.byref(c))
        return {"smem_bytes": a.value, "samples_per_cta": b.value, "threads_per_cta": c.value}

    # ------------------------------------------------------------------ device-level API (C ABI 1:1)
    def contact_stats(self, reset=True):
        """(sample-sub-steps with more candidate points inside the contact margin than max_contacts, sample-sub-steps
        simulated, sample-sub-steps with more TOUCHING points than max_contacts) since the last reset.  Synchronises."""
        out = (C.c_int64 * 3)()
        self._check(self._lib.sc_contact_stats(self._h, out, int(bool(reset))), "sc_contact_stats")
        return int(out[0]), int(out[1]), int(out[2])

    def empty_cache(self, n):
        """n contact-cache rows without contacts (ids -1, impulses 0)."""
        c = torch.zeros((n, CACHE_DIM), device=self.device, dtype=torch.float32)
        c[:, :CACHE_DIM // 2] = -1.0
        return c

    def window(self, parents, actions, targets, nsteps, children=None, parent_idx=None, sample_seq=None, out=None,
               parent_cache=None, out_cache=None, out_overflow=None):
        """One SamCon window for all samples.  Returns (state [S,132], cost [S], info [S,5]) on the device.
        ``parent_cache`` [E,CACHE_DIM]: the parents' contact caches (None: empty); ``out_cache`` [S,CACHE_DIM] receives
        every sample's cache at the end of the window (pass True to have it allocated; it is then a 4th return value);
        ``out_overflow`` int32 [S]: per sample, the sub-steps in which its candidate contacts exceeded max_contacts."""
        parents = self._f32(parents)
        actions = self._f32(actions)
        targets = self._f32(targets).reshape(-1, STATE_DIM)
        S, E, nseq = actions.shape[0], parents.shape[0], targets.shape[0]
        if actions.shape[1] != self.act_dim or parents.shape[1] != STATE_DIM:
            raise ValueError("bad parents/actions shape")
        parent_idx = self._i32(parent_idx)
        sample_seq = self._i32(sample_seq)
        if parent_idx is None:
            if children is None:
                if S % E:
                    raise ValueError("nSample should be a multiples of nSave")   # samcon.py:66-67
                children = S // E
        else:
            children = 0
        if out is None:
            out = (torch.empty((S, STATE_DIM), device=self.device, dtype=torch.float32),
                   torch.empty((S,), device=self.device, dtype=torch.float32),
                   torch.empty((S, 5), device=self.device, dtype=torch.float32))
        if parent_cache is not None:
            parent_cache = self._f32(parent_cache, (E, CACHE_DIM))
        want_cache = out_cache is True
        if want_cache:
            out_cache = torch.empty((S, CACHE_DIM), device=self.device, dtype=torch.float32)
        elif out_cache is not None and (tuple(out_cache.shape) != (S, CACHE_DIM) or out_cache.dtype != torch.float32
                                        or not out_cache.is_contiguous() or out_cache.device != self.device):
            raise ValueError("out_cache must be a contiguous float32 [S, CACHE_DIM] tensor on the pool's device")
        rc = self._lib.sc_window(self._h, _ptr(parents), _ptr(parent_cache), E, _ptr(parent_idx), int(children), _ptr(actions),
                                 _ptr(targets), nseq, _ptr(sample_seq), S, int(nsteps), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                 _ptr(out_cache), _ptr(out_overflow), self._stream())
        self._check(rc, "sc_window")
        return (*out, out_cache) if want_cache else out

    def step(self, state, actions, nsteps, cache=None):
        """In-place env.step (no cost) on [S,132] device states; ``cache`` [S,CACHE_DIM] is updated in place too."""
        state = self._f32(state)
        actions = self._f32(actions)
        if cache is not None and (tuple(cache.shape) != (state.shape[0], CACHE_DIM) or cache.dtype != torch.float32
                                  or not cache.is_contiguous() or cache.device != self.device):
            raise ValueError("cache must be a contiguous float32 [S, CACHE_DIM] tensor on the pool's device")
        self._check(self._lib.sc_step(self._h, _ptr(state), _ptr(cache), _ptr(actions), state.shape[0], int(nsteps),
                                      self._stream()), "sc_step")
        return state

    def cost(self, sim, kin):
        sim = self._f32(sim).reshape(-1, STATE_DIM)
        kin = self._f32(kin).reshape(-1, STATE_DIM)
        if kin.shape[0] == 1 and sim.shape[0] > 1:
            kin = kin.expand(sim.shape[0], STATE_DIM).contiguous()
        out = torch.empty((sim.shape[0], 6), device=self.device, dtype=torch.float32)
        self._check(self._lib.sc_cost(self._h, _ptr(sim), _ptr(kin), sim.shape[0], _ptr(out), self._stream()), "sc_cost")
        return out

    def sample_gen(self, target, limits, S, noise=None, perm=None, gaussian=False, seed=0, window=0, out=None):
        """SamCon.get_batch_action on the device.  ``target`` [132] (one sequence) or [Q,132]: Q sequences with S samples
        each in one launch, rows [q S, (q+1) S) drawn from the Philox stream (seed + q, window)."""
        target = self._f32(target).reshape(-1, STATE_DIM)
        Q = target.shape[0]
        limits = self._f32(limits, (3 * self.nj,))
        noise = self._f32(noise, (Q * S, 3 * self.nj)) if noise is not None else None
        perm = self._i32(perm)
        if out is None:
            out = torch.empty((Q * S, self.act_dim), device=self.device, dtype=torch.float32)
        elif tuple(out.shape) != (Q * S, self.act_dim) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise ValueError("out must be a contiguous float32 [Q*S, 4*nj] tensor on the pool's device")
        rc = self._lib.sc_sample_gen(self._h, _ptr(target), Q, _ptr(limits), _ptr(noise), _ptr(perm), int(bool(gaussian)),
                                     C.c_uint64(int(seed)), C.c_uint64(int(window)), int(S), _ptr(out), self._stream())
        self._check(rc, "sc_sample_gen")
        return out

    def motion_states(self, motions, times):
        """AmassMotionData.getSpecTimeState for all (sequence, time) pairs on the device.  ``motions``: list of
        (pose (T,75) array, key-frame duration) or AmassMotionData objects; ``times``: seconds.  -> [Q, len(times), 132]."""
        poses, offs, dts = [], [0], []
        for m in motions:
            pose, dur = (m._motion, m.keyFrameDuration()) if hasattr(m, "keyFrameDuration") else m
            pose = np.asarray(pose, dtype=np.float64)
            if pose.ndim != 2 or pose.shape[1] != 7 + 4 * self.nj:
                raise ValueError("pose arrays must be (T, 7 + 4*nj)")
            poses.append(pose); offs.append(offs[-1] + pose.shape[0]); dts.append(float(dur))
        d_pose = torch.tensor(np.concatenate(poses), dtype=torch.float32, device=self.device)
        d_off = torch.tensor(offs, dtype=torch.int32, device=self.device)
        d_dt = torch.tensor(dts, dtype=torch.float64, device=self.device)
        d_t = torch.tensor(np.asarray(times, dtype=np.float64).reshape(-1), dtype=torch.float64, device=self.device)
        out = torch.empty((len(poses), d_t.shape[0], STATE_DIM), device=self.device, dtype=torch.float32)
        self._check(self._lib.sc_motion_states(self._h, _ptr(d_pose), _ptr(d_off), _ptr(d_dt), len(poses), _ptr(d_t),
                                               int(d_t.shape[0]), _ptr(out), self._stream()), "sc_motion_states")
        return out

    def select(self, cost, E, seg_offsets=None):
        """Indices [nseg,E] and costs of the E smallest costs per segment, ascending (cost, index)."""
        cost = self._f32(cost).reshape(-1)
        if seg_offsets is None:
            seg_offsets = torch.tensor([0, cost.shape[0]], dtype=torch.int32, device=self.device)
        seg_offsets = self._i32(seg_offsets)
        nseg = seg_offsets.shape[0] - 1
        idx = torch.empty((nseg, E), device=self.device, dtype=torch.int32)
        ec = torch.empty((nseg, E), device=self.device, dtype=torch.float32)
        self._check(self._lib.sc_select(self._h, _ptr(cost), _ptr(seg_offsets), nseg, int(E), _ptr(idx), _ptr(ec),
                                        int(cost.shape[0]), self._stream()), "sc_select")
        return idx, ec

    def gather_rows(self, src, idx):
        src = self._f32(src)
        idx = self._i32(idx).reshape(-1)
        out = torch.empty((idx.shape[0], src.shape[1]), device=self.device, dtype=torch.float32)
        self._check(self._lib.sc_gather_rows(self._h, _ptr(src), _ptr(idx), idx.shape[0], src.shape[1], _ptr(out),
                                             self._stream()), "sc_gather_rows")
        return out

    def gather_rows_peer(self, peer_base, src_peer, src_row, width):
        """Rows out of other GPUs' buffers (sc_gather_rows_peer): peer_base int64 [npeer] device addresses of [*,width]
        float32 arrays mapped into this process, src_peer / src_row int32 [n]."""
        src_peer, src_row = self._i32(src_peer).reshape(-1), self._i32(src_row).reshape(-1)
        if peer_base.dtype != torch.int64 or peer_base.device != self.device:
            raise ValueError("peer_base must be an int64 tensor on the pool's device")
        out = torch.empty((src_row.shape[0], int(width)), device=self.device, dtype=torch.float32)
        self._check(self._lib.sc_gather_rows_peer(self._h, _ptr(peer_base), _ptr(src_peer), _ptr(src_row), src_row.shape[0],
                                                  int(width), _ptr(out), self._stream()), "sc_gather_rows_peer")
        return out

    # ------------------------------------------------------------------ worker-protocol mirror (host arrays)
    def init(self, init_state, nSave=1):
        """("init", [iter0, uid0, init_state]) -- samcon.py:42-46: every elite slot starts from init_state."""
        s = self._f32(np.asarray(init_state, dtype=np.float64), (STATE_DIM,))
        self._elites = s.reshape(1, STATE_DIM).repeat(int(nSave), 1).contiguous()
        self._elite_cache = None

    def set_elites(self, states, caches=None):
        self._elites = self._f32(states).reshape(-1, STATE_DIM).contiguous()
        self._elite_cache = self._f32(caches).reshape(-1, CACHE_DIM).contiguous() if caches is not None else None

    @property
    def elites(self):
        return self._elites

    def sim(self, target_state, actions, nsteps):
        """("sim", ...) for all jobs at once -- samcon.py:28-41.  Host in, host out (float64 like the reference)."""
        if self._elites is None:
            raise RuntimeError("init() must be called before sim()")
        if not isinstance(actions, torch.Tensor):
            actions = np.asarray(actions, dtype=np.float32)
        S = len(actions)
        if getattr(self, "_ovf", None) is None or self._ovf.shape[0] != S:
            self._ovf = torch.empty((S,), dtype=torch.int32, device=self.device)
        st, cost, info, cache = self.window(self._elites, actions, np.asarray(target_state, dtype=np.float32), nsteps,
                                            parent_cache=self._elite_cache, out_cache=True, out_overflow=self._ovf)
        self._last = (st, cost, info, cache)
        # device -> pinned host staging (allocated once per sample count), one synchronisation for the three results;
        # the widening to the reference's float64 happens on the device (a numpy astype of 1.4 M values costs more
        # than the extra 5.5 MB over PCIe)
        S = st.shape[0]
        if getattr(self, "_stage_S", None) != S:
            self._stage = tuple(torch.empty(t.shape, dtype=torch.float64, pin_memory=True) for t in (st, cost, info))
            self._stage_S = S
        for h, d in zip(self._stage, (st, cost, info)):
            h.copy_(d.double(), non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return tuple(h.numpy().copy() for h in self._stage)

    def select_last(self, nSave):
        """Trajectory.select('greedy') on the device for the window just simulated -> host int64 indices."""
        idx, _ = self.select(self._last[1], nSave)
        return idx[0].cpu().numpy().astype(np.int64)

    def keep(self, elite_inds):
        """Make the samples picked by Trajectory.select the parents of the next window (state stays in HBM)."""
        idx = self._i32(np.asarray(elite_inds, dtype=np.int32))
        self._elites = self.gather_rows(self._last[0], idx)
        self._elite_cache = self.gather_rows(self._last[3], idx)      # the snapshot carries the contact cache
        self.elite_overflow = self._ovf[idx.long()]                   # per kept sample: sub-steps over the contact cap

    def close(self):
        if not self._closed and self._h:
            torch.cuda.synchronize(self.device)
            self._lib.sc_destroy(self._h)
            self._h = None
            self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""The SamCon search loop over the GPU sample pool.

Same constructor, cfg keys and loop shape as the reference's ``SamCon``
(/root/reference/samcon/core/samcon.py:54-198): per window draw nSample perturbed PD targets around the
reference pose, roll every sample forward from one of the nSave elites of the previous window, score it, keep
the nSave best.  What changed is below the pipe: instead of scattering (prev_uids, actions) to
``num_processes`` PyBullet workers and gathering pickled result lists, one ``GpuSamplePool.sim`` call runs all
samples on the GPU and elite states never leave HBM (``pool.keep``).
"""
from __future__ import annotations

import datetime
import logging
import os
import time

import numpy as np

from .model import Character
from .sampling import draw_noise, get_batch_action
from .trajectory import Trajectory


class UniqueId:
    """/root/reference/samcon/core/unique_id.py:1-12"""

    def __init__(self):
        self._id = 0

    def get(self, num=1):
        ids = range(self._id, self._id + num)
        self._id += num
        return iter(ids)

    def reset(self):
        self._id = 0


def get_eta_str(cur_iter, total_iter, time_per_iter):     # /root/reference/utils/tools.py:17-19
    eta = (total_iter - cur_iter) * time_per_iter
    return str(datetime.timedelta(seconds=round(eta)))


class SamCon:

    def __init__(self, cfg, data_loader, num_processes=1, nIter=None, nSample=None, nSave=None, pool=None, logger=None,
                 device_select=True, device_actions=True):
        """``num_processes`` is accepted for call compatibility (run_samcon.py:36); the GPU pool replaces the
        process pool, so its value is not used.  ``pool`` may be any object with init/sim/keep/close (the
        tests pass an oracle-backed one); by default a GpuSamplePool on cuda:0 is created."""
        self._rng = np.random.RandomState(cfg.seed)        # np.random.seed(cfg.seed), samcon.py:57
        self._cfg = cfg
        self._data_loader = data_loader
        self._sample_timestep = 1.0 / cfg.fps_act
        self._nIter = cfg.nIter if nIter is None else nIter
        self._nSample = cfg.nSample if nSample is None else nSample
        self._nSave = cfg.nSave if nSave is None else nSave
        if self._nSample % self._nSave != 0:
            raise Exception("nSample should be a multiples of nSave")
        self._num_substep = self._nSample // self._nSave
        if cfg.fps_sim % cfg.fps_act != 0:
            raise Exception("FPS_SIM should be a multiples of FPS_ACT")      # humanoid_tracking_env.py:27-28
        self._steps_per_window = cfg.fps_sim // cfg.fps_act
        self._num_processes = num_processes
        if pool is None:
            from .pool import GpuSamplePool
            pool = GpuSamplePool(Character.from_config(cfg), 0)
        self._pool = pool
        self._device_select = device_select and hasattr(pool, "select_last")
        self._device_actions = device_actions and hasattr(pool, "sample_gen") and cfg.noise_type == "uniform"
        self._unique_id = UniqueId()
        self._traj = Trajectory(self._nIter, self._nSample, self._nSave)
        self._logger = logger or logging.getLogger("samcon_b200")
        self.window_times = []

    def get_batch_action(self, state):
        return get_batch_action(state, self._nSample, self._cfg.values, self._cfg.noise_type, self._rng)

    def sample(self, save=True):
        startp_iter = 0
        startp_uid = next(self._unique_id.get(1))
        startp_state = self._data_loader.getSpecTimeState(t=0.0)
        self._pool.init(startp_state, self._nSave)
        for _ in range(self._nSample):
            self._traj.push(startp_iter, [-1, startp_uid, startp_state, None, None, 0.0, None])
        self._traj.select(startp_iter)

        for t in range(1, self._nIter + 1):
            t0 = time.time()
            target_state = self._data_loader.getSpecTimeState(t * self._sample_timestep)
            if self._device_actions:
                # same MT19937 draws and shuffle as get_batch_action (samcon.py:113-118, 140); the Euler <-> quaternion
                # arithmetic (samcon.py:121-137) runs on the device (sc_sample_gen) instead of scipy on the host
                noise, perm = draw_noise(self._nSample, self._cfg.values, self._cfg.noise_type, self._rng)
                dev_action = self._pool.sample_gen(np.asarray(target_state, dtype=np.float32),
                                                   np.asarray(self._cfg.values, dtype=np.float32), self._nSample,
                                                   noise=noise.astype(np.float32), perm=perm.astype(np.int32))
                batch_action = None
            else:
                batch_action = self.get_batch_action(target_state)
                dev_action = batch_action
            prev_uids = self._traj.get_curr_uid_elite(t - 1)
            curr_uids = np.fromiter(self._unique_id.get(self._nSample), dtype=np.int64, count=self._nSample)
            # sample k descends from elite k // num_substep (samcon.py:31-35, 85-90)
            sim_state, cost, info = self._pool.sim(target_state, dev_action, self._steps_per_window)
            if batch_action is None:
                batch_action = dev_action.cpu().numpy().astype(np.float64)
            self._traj.push_batch(t, np.repeat(prev_uids, self._num_substep), curr_uids, target_state, sim_state,
                                  batch_action, cost, info)
            elite = self._pool.select_last(self._nSave) if self._device_select else None
            self._traj.select(t, "greedy", elite)
            self._pool.keep(self._traj.get_elite_inds(t))
            t1 = time.time()
            self.window_times.append(t1 - t0)
            self._logger.info("cfg: {} | iter {}/{} | T_i {:.2f}s | ETA {}".format(
                self._cfg.id, t, self._nIter, t1 - t0, get_eta_str(t, self._nIter, t1 - t0)))

        if save:
            os.makedirs(self._cfg.output_dir, exist_ok=True)
            os.makedirs(self._cfg.info_dir, exist_ok=True)
            res = self._traj.find_path_and_save(self._cfg)
        else:
            res = self._traj.build_results(self._cfg.motion_seq)
        self._logger.info(f'sampling done {datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S")}')
        return res

    def close(self):
        self._pool.close()

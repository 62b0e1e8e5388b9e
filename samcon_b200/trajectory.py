"""Per-window sample store, elite selection and best-path extraction.

Same interface and the same result/info pickle layout as the reference's ``Trajectory``
(/root/reference/samcon/core/trajectory.py:7-201), restated without ``treelib`` (a uid -> parent-uid map
does the back-trace) and without the removed ``np.int`` alias.  ``push_batch`` is the array-at-a-time entry
the GPU pool uses; ``push`` keeps the reference's one-sample signature.
"""
from __future__ import annotations

import datetime
import os.path as osp

import numpy as np


class Trajectory:

    def __init__(self, nIter, nSample, nSave):
        self._parent_of = {0: None}         # uid tree (root uid 0), trajectory.py:16-17
        self._nIter = nIter
        self._nSample = nSample
        self._nSave = nSave
        n = nIter + 1
        self.prev_uid = [[] for _ in range(n)]
        self.curr_uid = [[] for _ in range(n)]
        self.target_state = [[] for _ in range(n)]
        self.simulated_state = [[] for _ in range(n)]
        self.action = [[] for _ in range(n)]
        self.cost = [[] for _ in range(n)]
        self.info = [[] for _ in range(n)]
        self.elite_inds = [[] for _ in range(n)]

    def push(self, i, data):
        d1, d2, d3, d4, d5, d6, d7 = data
        self.prev_uid[i].append(d1)
        self.curr_uid[i].append(d2)
        self.target_state[i].append(d3)
        self.simulated_state[i].append(d4)
        self.action[i].append(d5)
        self.cost[i].append(d6)
        self.info[i].append(d7)
        if i > 0:
            self._parent_of[int(d2)] = int(d1)

    def push_batch(self, i, prev_uid, curr_uid, target_state, simulated_state, action, cost, info):
        n = len(curr_uid)
        self.prev_uid[i] = list(np.asarray(prev_uid).tolist())
        self.curr_uid[i] = list(np.asarray(curr_uid).tolist())
        self.target_state[i] = [target_state] * n
        self.simulated_state[i] = simulated_state
        self.action[i] = action
        self.cost[i] = cost
        self.info[i] = info
        if i > 0:
            self._parent_of.update(zip(self.curr_uid[i], self.prev_uid[i]))

    def get_elite_inds(self, i):
        return np.array(self.elite_inds[i], dtype=np.int64)

    def _column(self, field, i, dtype):
        return np.array(getattr(self, field)[i], dtype=dtype)

    def select(self, i, mode="greedy", elite_inds=None):
        """trajectory.py:108-116.  ``elite_inds`` lets the caller hand in the device top-k result (same order:
        ascending cost, ties by index -- the reference's argsort leaves ties unspecified)."""
        assert len(self.cost[i]) == self._nSample
        if mode == "greedy":
            if elite_inds is None:
                c = self.get_cost(i)
                elite_inds = np.lexsort((np.arange(len(c)), c))[:self._nSave]
            self.elite_inds[i] = np.asarray(elite_inds, dtype=np.int64)
        elif mode == "diversity":
            raise NotImplementedError
        else:
            raise NotImplementedError

    def _rsearch(self, uid):
        out = []
        while uid is not None:
            out.append(uid)
            uid = self._parent_of[uid]
        return out

    def find_paths(self):
        """-> (paths_my_sys (nSave, nIter+1) elite-slot indices, total_cost (nSave,), order)"""
        end_points = self.get_curr_uid_elite(-1)
        paths = np.array([self._rsearch(int(ep))[::-1] for ep in end_points], dtype=np.int64)
        num_path = paths.shape[0]
        paths_my_sys = paths.copy()
        for i in range(1, self._nIter + 1):
            candidates = self.get_curr_uid_elite(i)
            pos = {int(u): k for k, u in enumerate(candidates)}
            paths_my_sys[:, i] = [pos[int(paths[j, i])] for j in range(num_path)]
        total_cost = np.zeros(num_path)
        for i in range(1, self._nIter + 1):
            total_cost += self.get_cost_elite(i)[paths_my_sys[:, i]]
        return paths_my_sys, total_cost, total_cost.argsort(kind="stable")

    def build_results(self, motion_seq):
        """The two dicts the reference pickles (trajectory.py:151-197)."""
        paths_my_sys, total_cost, order = self.find_paths()
        elite = {t: (self.get_target_state_elite(t), self.get_simulated_state_elite(t), self.get_action_elite(t),
                     self.get_cost_elite(t)) for t in range(1, self._nIter + 1)}
        out = {motion_seq: {}}
        t0 = self.get_target_state_elite(0)[0]
        for i in range(paths_my_sys.shape[0]):
            single_path = paths_my_sys[order[i]]
            data = {"target_state": [], "simulated_state": [], "action": [], "cost": []}
            for t in range(1, self._nIter + 1):
                k = single_path[t]
                ts, ss, ac, co = elite[t]
                data["target_state"].append(ts[k])
                data["simulated_state"].append(ss[k])
                data["action"].append(ac[k])
                data["cost"].append(co[k])
            data = {k: np.vstack(v) for k, v in data.items()}
            data["total_cost"] = np.sum(data["cost"])
            data["t0_state"] = t0
            out[motion_seq][f"path_{i}"] = data
        info = {k: [] for k in ("pose_cost", "root_cost", "ee_cost", "balance_cost", "com_cost", "total_cost", "elite_inds")}
        for t in range(1, self._nIter + 1):
            d = self.get_info(t)
            for c, k in enumerate(("pose_cost", "root_cost", "ee_cost", "balance_cost", "com_cost")):
                info[k].append(d[:, c])
            info["total_cost"].append(self.get_cost(t))
            info["elite_inds"].append(self.get_elite_inds(t))
        info = {k: np.vstack(v) for k, v in info.items()}
        info["best_path_my_sys"] = paths_my_sys[order[0]][1:]
        return out, info

    def find_path_and_save(self, cfg):
        import joblib
        out, info = self.build_results(cfg.motion_seq)
        time_flag = datetime.datetime.now().strftime("%Y-%m-%d-%H-%M-%S")
        joblib.dump(out, osp.join(cfg.output_dir, f"{cfg.id}_{time_flag}.pkl"))
        joblib.dump(info, osp.join(cfg.info_dir, f"{cfg.id}_{time_flag}.pkl"))
        return out, info


# The reference's accessor family (trajectory.py:57-106): get_<field>(i) for all nSample samples of window i and
# get_<field>_elite(i) for its nSave elites, generated from the field table instead of written out fourteen times.
def _install_getters():
    fields = {"prev_uid": np.int64, "curr_uid": np.int64, "target_state": np.float64, "simulated_state": np.float64,
              "action": np.float64, "cost": np.float64, "info": np.float64}
    for field, dtype in fields.items():
        def get_all(self, i, _f=field, _d=dtype):
            return self._column(_f, i, _d)

        def get_elite(self, i, _f=field, _d=dtype):
            return self._column(_f, i, _d)[self.get_elite_inds(i)]
        get_all.__name__, get_elite.__name__ = f"get_{field}", f"get_{field}_elite"
        setattr(Trajectory, get_all.__name__, get_all)
        setattr(Trajectory, get_elite.__name__, get_elite)


_install_getters()

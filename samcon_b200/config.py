"""SamCon configuration: same yml keys and attribute names as the reference's
``samcon.utils.config.Config`` (/root/reference/samcon/utils/config.py:8-70), so search-loop
code written against the reference reads a cfg the same way.

Differences, all on the host and outside the hot path:
* cfg files are looked up in ``<base_dir>/samcon/cfg`` (reference layout) and then in the
  package's bundled ``data/cfg``;
* directories are created lazily (``create_dirs=True``) and never wiped -- the GPU path keeps
  sample states in HBM, so there is no ``states`` spill directory to recreate.
"""
from __future__ import annotations

import os
import os.path as osp

import numpy as np
import yaml

_PKG_CFG = osp.join(osp.dirname(osp.abspath(__file__)), "data", "cfg")
_SCALAR_KEYS = ("fps_sim", "fps_act", "seed", "auto_nIter", "nIter", "nSample", "nSave", "data_path", "motion_seq", "noise_type",
                "model_file", "scale", "height", "self_collision", "joint_weights", "end_effectors", "cost_weights")


def _column(rows, c):
    return np.array([float(r[c]) for r in rows])


class Config:

    def __init__(self, cfg_id, base_dir="", create_dirs=False, cfg_dict=None):
        self.id = cfg_id
        base_dir = base_dir if base_dir else ""
        self.base_dir = osp.expanduser(base_dir)

        if cfg_dict is not None:
            cfg = cfg_dict
        else:
            cands = [osp.join(self.base_dir, "samcon", "cfg", f"{cfg_id}.yml"), osp.join(_PKG_CFG, f"{cfg_id}.yml")]
            path = next((c for c in cands if osp.isfile(c)), None)
            if path is None:
                raise FileNotFoundError(f"no cfg '{cfg_id}' in {cands}")
            with open(path, "rb") as f:
                cfg = yaml.safe_load(f)
        self.cfg_dict = cfg

        self.result_dir = osp.join(self.base_dir, "results")
        self.cfg_dir = "%s/samcon/%s" % (self.result_dir, cfg_id)
        self.output_dir = "%s/results" % self.cfg_dir
        self.log_dir = "%s/log" % self.cfg_dir
        self.info_dir = "%s/info" % self.cfg_dir
        self.state_dir = "%s/states" % cfg.get("temp_dir", "/tmp/samcon_b200")
        if create_dirs:
            for d in (self.output_dir, self.info_dir, self.log_dir):
                os.makedirs(d, exist_ok=True)

        # cfg keys that become attributes of the same name (reference: samcon/utils/config.py:34-68)
        for key in _SCALAR_KEYS:
            setattr(self, key, cfg[key])
        # per-joint tables: rows of [name, value, ...] in the yml, columns as arrays here
        if "joint_noises" in cfg:
            self.values = _column(cfg["joint_noises"], 1)                      # noise range / variance per Euler angle
        if "joint_params" in cfg:
            self.kps, self.kds, self.max_forces = (_column(cfg["joint_params"], c) for c in (1, 2, 3))
            self.char_info = {"kps": self.kps, "kds": self.kds, "max_forces": self.max_forces}

    def get(self, key, default=None):
        return self.cfg_dict.get(key, default)

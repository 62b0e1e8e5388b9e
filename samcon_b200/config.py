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

        self.fps_sim = cfg["fps_sim"]
        self.fps_act = cfg["fps_act"]
        self.seed = cfg["seed"]
        self.auto_nIter = cfg["auto_nIter"]
        self.nIter = cfg["nIter"]
        self.nSample = cfg["nSample"]
        self.nSave = cfg["nSave"]

        self.data_path = cfg["data_path"]
        self.motion_seq = cfg["motion_seq"]

        self.noise_type = cfg["noise_type"]
        if "joint_noises" in cfg:
            self.values = np.array([float(r[1]) for r in cfg["joint_noises"]])

        self.model_file = cfg["model_file"]
        self.scale = cfg["scale"]
        self.height = cfg["height"]
        self.self_collision = cfg["self_collision"]
        if "joint_params" in cfg:
            self.kps = np.array([float(r[1]) for r in cfg["joint_params"]])
            self.kds = np.array([float(r[2]) for r in cfg["joint_params"]])
            self.max_forces = np.array([float(r[3]) for r in cfg["joint_params"]])
            self.char_info = {"kps": self.kps, "kds": self.kds, "max_forces": self.max_forces}
        self.joint_weights = cfg["joint_weights"]
        self.end_effectors = cfg["end_effectors"]
        self.cost_weights = cfg["cost_weights"]

    def get(self, key, default=None):
        return self.cfg_dict.get(key, default)

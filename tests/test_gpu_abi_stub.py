"""INTEGRATION.md §3's raw ctypes stub, executed verbatim on the GPU: the code block is cut
out of the document and exec'd with the reference-side names it expects (`model`, `cfg`, `loader`, `init_state`,
`samcon`) in scope; its elites must be the ones the maintained host loop (SamCon.sample over GpuSamplePool) keeps."""
import copy
import ctypes as C
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_stub_runs_verbatim(walk_cfg, character):
    import torch
    from samcon_b200 import abi
    from samcon_b200.motion import AmassMotionData, make_synthetic_motion
    from samcon_b200.pool import GpuSamplePool
    from samcon_b200.samcon import SamCon
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"## 3\. Raw `ctypes` stub.*?```python\n(.*?)```", text, flags=re.S).group(1)
    m = make_synthetic_motion("gait", T=40, seed=0)
    (name, _), = m.items()
    loader = AmassMotionData(m, name)
    cfg = copy.copy(walk_cfg)
    cfg.nSample, cfg.nSave, cfg.nIter, cfg.motion_seq = 200, 20, 4, name
    model, keep = abi.make_model(character)

    class _Host:                      # stands in for the reference's SamCon object: only get_batch_action is used
        def __init__(self):
            self._rng = np.random.RandomState(cfg.seed)

        def get_batch_action(self, state):
            from samcon_b200.sampling import get_batch_action
            return get_batch_action(state, cfg.nSample, cfg.values, cfg.noise_type, self._rng)
    cwd = os.getcwd()
    os.chdir(ROOT)                    # the stub opens the library by its repo-relative path
    try:
        # the block, verbatim, in two pieces: its declarations first, because `model` has to be an instance of the struct
        # class the block itself declares (same bytes as the maintained abi.ScModel)
        cut = block.index("# `model` = sc_model(...)")
        ns = {"cfg": cfg, "loader": loader, "init_state": loader.getSpecTimeState(0.0), "samcon": _Host()}
        exec(compile(block[:cut], "INTEGRATION.md#3a", "exec"), ns)
        assert C.sizeof(ns["sc_model"]) == C.sizeof(abi.ScModel)
        ns["model"] = ns["sc_model"].from_buffer_copy(bytes(model))
        exec(compile(block[cut:], "INTEGRATION.md#3b", "exec"), ns)
    finally:
        os.chdir(cwd)
    stub_elites = ns["elites"].cpu().numpy()
    pool = GpuSamplePool(character, 0)
    algo = SamCon(cfg, loader, 1, cfg.nIter, cfg.nSample, cfg.nSave, pool=pool, device_actions=False)
    try:
        algo.sample(save=False)
        np.testing.assert_array_equal(stub_elites, pool.elites.cpu().numpy())
    finally:
        algo.close()
    assert torch.isfinite(ns["ec"]).all()

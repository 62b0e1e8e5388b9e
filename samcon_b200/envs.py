"""Single-sample env facade with the reference's env interface, for replay / evaluation.

``scripts/eval_samcon.py`` (/root/reference/scripts/eval_samcon.py:24, 67-81) drives
``HumanoidTrackingEnv(cfg, viz)``: ``env._sim_agent.set_state(t0)``, then per window
``env._kin_agent.set_state(target)``; ``env.step(action)``.  This class keeps those names and return values
(/root/reference/envs/humanoid_tracking_env.py:184-200, /root/reference/envs/humanoid_stable_pd.py:137-165)
on top of the GPU pool with S = 1.  There is no GUI (``viz`` must be False).
"""
from __future__ import annotations

import numpy as np

from .model import STATE_DIM, Character


class _Agent:
    """State holder with HumanoidStablePD's set_state/get_state and pQvw accessors."""

    def __init__(self, nj, on_set=None):
        self._nj = nj
        self._on_set = on_set
        self._state = np.zeros(STATE_DIM)
        self._state[6] = 1.0
        self._state[10:7 + 4 * nj:4] = 1.0
        self._joint_indices_movable = list(range(nj))

    def set_state(self, state):
        state = np.array(state, dtype=np.float64)
        assert state.shape == (STATE_DIM,)
        self._state = state
        if self._on_set:
            self._on_set()

    def get_state(self):
        return self._state.copy()

    def get_base_pQvw(self):
        s, n = self._state, 7 + 4 * self._nj
        return s[0:3].copy(), s[3:7].copy(), s[n:n + 3].copy(), s[n + 3:n + 6].copy()

    def get_joint_pv(self, indices=None):
        s, n = self._state, 7 + 4 * self._nj
        ps = s[7:n].reshape(-1, 4)
        vs = s[n + 6:].reshape(-1, 3)
        if indices is None:
            indices = range(self._nj)
        indices = list(indices)
        if len(indices) == 1:
            return ps[indices[0]].copy(), vs[indices[0]].copy()
        return [ps[i].copy() for i in indices], [vs[i].copy() for i in indices]


class HumanoidTrackingEnv:

    def __init__(self, cfg, viz=False, pool=None):
        if viz:
            raise NotImplementedError("the GPU path is headless")
        self._cfg = cfg
        self._dt_sim = 1.0 / cfg.fps_sim
        self._dt_act = 1.0 / cfg.fps_act
        if cfg.fps_sim % cfg.fps_act != 0:
            raise Exception("FPS_SIM should be a multiples of FPS_ACT")
        self._num_substep = cfg.fps_sim // cfg.fps_act
        if pool is None:
            from .pool import GpuSamplePool
            pool = GpuSamplePool(Character.from_config(cfg), 0)
        self._pool = pool
        nj = (STATE_DIM - 13) // 7
        # the world's contact cache lives between step() calls like Bullet's manifolds do; teleporting the simulated
        # character (set_state) drops it
        self._cache = None
        self._sim_agent = _Agent(nj, self._drop_cache)
        self._kin_agent = _Agent(nj)

    def _drop_cache(self):
        self._cache = None

    def compute_total_cost(self):
        c = self._pool.cost(self._sim_agent.get_state().astype(np.float32)[None],
                            self._kin_agent.get_state().astype(np.float32)[None]).cpu().numpy()[0].astype(np.float64)
        return tuple(c)

    def step(self, action):
        action = np.asarray(action, dtype=np.float32).reshape(1, -1)
        st, cost, info, cache = self._pool.window(self._sim_agent.get_state().astype(np.float32)[None], action,
                                                  self._kin_agent.get_state().astype(np.float32), self._num_substep,
                                                  children=1, parent_cache=self._cache, out_cache=True)
        simulated_state = st.cpu().numpy()[0].astype(np.float64)
        self._sim_agent.set_state(simulated_state)
        self._cache = cache
        return simulated_state, float(cost.cpu().numpy()[0]), [float(x) for x in info.cpu().numpy()[0]]

    def close(self):
        self._pool.close()

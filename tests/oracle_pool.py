"""Oracle-backed stand-in for GpuSamplePool (tests only): same init/sim/keep/close protocol on the CPU oracle, so the
host search loop can be exercised without a GPU."""
import numpy as np


class OraclePool:
    def __init__(self, oracle):
        self.o = oracle
        self._elites = None
        self._last = None

    def init(self, init_state, nSave=1):
        self._elites = np.repeat(np.asarray(init_state, dtype=np.float64)[None], nSave, axis=0)

    def sim(self, target_state, actions, nsteps):
        st, cost, info = self.o.window(self._elites, np.asarray(actions, np.float64), np.asarray(target_state, np.float64), nsteps)
        self._last = (st, cost, info)
        return st, cost, info

    def keep(self, elite_inds):
        self._elites = self._last[0][np.asarray(elite_inds)]

    def close(self):
        pass

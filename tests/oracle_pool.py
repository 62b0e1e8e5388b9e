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
        self._cache = None

    def sim(self, target_state, actions, nsteps):
        st, cost, info, cache = self.o.window(self._elites, np.asarray(actions, np.float64), np.asarray(target_state, np.float64),
                                              nsteps, parent_cache=self._cache, return_cache=True)
        self._last = (st, cost, info, cache)
        return st, cost, info

    def keep(self, elite_inds):
        self._elites = self._last[0][np.asarray(elite_inds)]
        self._cache = self._last[3][np.asarray(elite_inds)]

    def close(self):
        pass


class OracleDevicePool:
    """CPU stand-in for GpuSamplePool's device-level API (window / select / gather_rows / sample_gen on torch CPU
    tensors, arithmetic by the oracle and the host sampler): lets the batched multi-sequence loop run without a GPU."""

    def __init__(self, oracle, cfg):
        import torch
        self.o, self.cfg, self.torch = oracle, cfg, torch
        self.device = torch.device("cpu")
        self.nj = oracle.nj
        self.act_dim = 4 * self.nj

    def sample_gen(self, target, limits, S, noise=None, perm=None, gaussian=False, seed=0, window=0, out=None):
        from samcon_b200.sampling import get_batch_action
        tg = np.asarray(target, np.float64).reshape(-1, 132)
        rows = []
        for q in range(len(tg)):                      # sequence q draws from the stream (seed + q, window), like the device
            rs = np.random.RandomState(((int(seed) + q) * 1000003 + int(window)) % (2 ** 31))
            rows.append(get_batch_action(tg[q], S, np.asarray(limits, np.float64), "gaussian" if gaussian else "uniform", rs))
        a = np.concatenate(rows)
        t = self.torch.tensor(a, dtype=self.torch.float32)
        if out is not None:
            out.copy_(t)
            return out
        return t

    def window(self, parents, actions, targets, nsteps, children=None, parent_idx=None, sample_seq=None, out=None,
               parent_cache=None, out_cache=None):
        P = np.asarray(parents, np.float64); A = np.asarray(actions, np.float64)
        T = np.asarray(targets, np.float64).reshape(-1, 132)
        S = len(A)
        pidx = np.asarray(parent_idx) if parent_idx is not None else np.arange(S) // (children or S // len(P))
        seq = np.asarray(sample_seq) if sample_seq is not None else np.zeros(S, int)
        st = np.zeros((S, 132)); cost = np.zeros(S); info = np.zeros((S, 5)); cache = np.zeros((S, self.o.cache_dim))
        PC = np.asarray(parent_cache, np.float64) if parent_cache is not None else None
        for k in range(S):
            s, c, i, cc = self.o.window(P[pidx[k]][None], A[k][None], T[seq[k]], nsteps,
                                        parent_cache=PC[pidx[k]][None] if PC is not None else None, return_cache=True)
            st[k], cost[k], info[k], cache[k] = s[0], c[0], i[0], cc[0]
        f = lambda x: self.torch.tensor(x, dtype=self.torch.float32)
        res = (f(st), f(cost), f(info))
        if out is not None:                            # like the GPU pool: results land in the caller's tensors
            for dst, src in zip(out, res):
                dst.copy_(src)
            res = tuple(out)
        if out_cache is True:
            return (*res, f(cache))
        if out_cache is not None:
            out_cache.copy_(f(cache))
        return res

    def select(self, cost, E, seg_offsets=None):
        c = np.asarray(cost, np.float64).reshape(-1)
        seg = np.asarray(seg_offsets) if seg_offsets is not None else np.array([0, len(c)])
        idx = np.zeros((len(seg) - 1, E), np.int32); ec = np.zeros((len(seg) - 1, E), np.float32)
        for i in range(len(seg) - 1):
            cs = c[seg[i]:seg[i + 1]]
            o = np.lexsort((np.arange(len(cs)), cs))[:E]
            idx[i] = seg[i] + o; ec[i] = cs[o]
        return self.torch.tensor(idx), self.torch.tensor(ec)

    def gather_rows(self, src, idx):
        return src[self.torch.as_tensor(idx).reshape(-1).long()].contiguous()

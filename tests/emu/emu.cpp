// Host emulation of the rollout kernel's lane stages (samcon_b200/csrc/sc_core.cuh compiled with g++).
// DEBUGGING / CPU-TEST AID ONLY: lets `pytest -m "not gpu"` exercise the kernel's arithmetic against the
// oracle on a box without a GPU.  Never loaded by the samcon_b200 package; not a fallback.
#include <stdlib.h>
#include <vector>
#include "../../samcon_b200/csrc/sc_tables.h"

extern "C" {
int emu_window(const sc_model *m, const float *parents, int E, const int *parent_idx, int children,
               const float *actions, const float *targets, int nseq, const int *sample_seq, int S, int nsteps,
               float *out_state, float *out_cost, float *out_info, const float *parent_cache, float *out_cache, char *err,
               int errlen) {
    sc::TablesBox T;          // (the emulation always carries the exact box narrow phase; nbox = 0 switches it off at run time)
    int rc = sc::build_tables(m, &T, err, errlen);
    if (rc) return rc;
    std::vector<float> sm(sc::WORDS_PER_TILE), tfeat((size_t)nseq * 18);
    // target features: the targets themselves pushed through the feature pass
    sc::RolloutArgs f = {};
    f.parents = targets; f.children = 1; f.S = nseq; f.nsteps = 0; f.out_feat = tfeat.data();
    for (int t0 = 0; t0 < nseq; t0 += sc::TILE) sc::rollout_tile(sm.data(), T, f, t0, 0, 32, nullptr, 0);
    sc::RolloutArgs a = {};
    a.parents = parents; a.parent_idx = parent_idx; a.children = children; a.actions = actions; a.targets = targets;
    a.tfeat = tfeat.data(); a.sample_seq = sample_seq; a.S = S; a.nsteps = nsteps;
    a.out_state = out_state; a.out_cost = out_cost; a.out_info = out_info;
    a.parent_cache = parent_cache; a.out_cache = out_cache;
    for (int t0 = 0; t0 < S; t0 += sc::TILE) sc::rollout_tile(sm.data(), T, a, t0, 0, 32, nullptr, 0);
    return 0;
}
int emu_words_per_sample() { return sc::WORDS_PER_SAMPLE; }
int emu_tables_bytes() { return (int)sizeof(sc::Tables); }
}

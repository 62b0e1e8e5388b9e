// The rollout kernel's source (samcon_b200/csrc/sc_core.cuh) compiled for the host with every `float` turned into `double`:
// the same algorithms (articulated-body recursions, contact-space Gauss-Seidel, same stage order) at the oracle's
// precision.  TEST AID ONLY (tests/test_f64_kernel_build.py): it separates what float32 rounding does to a rollout from
// what the kernel's algorithms do -- on chaotic contact-rich samples the double build leaves the oracle exactly where the
// float build does.  Never loaded by the samcon_b200 package.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../include/samcon_b200.h"      // the C ABI keeps its float32 arrays: included before the substitution
#define float double
#define sqrtf sqrt
#define fabsf fabs
#define atan2f atan2
#define fminf fmin
#define fmaxf fmax
#define sinf sin
#define cosf cos
#include "../../samcon_b200/csrc/sc_tables.h"
#undef float

extern "C" {
int emu64_window(const sc_model *m, const double *parents, int E, int children, const double *actions, const double *targets,
                 int nseq, int S, int nsteps, double *out_state, double *out_cost, double *out_info, char *err, int errlen) {
    sc::TablesBox T;          // (the emulation always carries the exact box narrow phase; nbox = 0 switches it off at run time)
    int rc = sc::build_tables(m, &T, err, errlen);
    if (rc) return rc;
    std::vector<double> sm(sc::WORDS_PER_TILE), tfeat((size_t)nseq * 18);
    sc::RolloutArgs f = {};
    f.parents = targets; f.children = 1; f.S = nseq; f.nsteps = 0; f.out_feat = tfeat.data();
    for (int t0 = 0; t0 < nseq; t0 += sc::TILE) sc::rollout_tile(sm.data(), T, f, t0, 0, 32, nullptr, 0);
    sc::RolloutArgs a = {};
    a.parents = parents; a.children = children; a.actions = actions; a.targets = targets;
    a.tfeat = tfeat.data(); a.S = S; a.nsteps = nsteps;
    a.out_state = out_state; a.out_cost = out_cost; a.out_info = out_info;
    for (int t0 = 0; t0 < S; t0 += sc::TILE) sc::rollout_tile(sm.data(), T, a, t0, 0, 32, nullptr, 0);
    return 0;
}
}

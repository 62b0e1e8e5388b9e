// sc_tables.h -- host: derive the kernel's lookup tables (sc::Tables) from the C-ABI model description.
#pragma once
#include <stdio.h>
#include <math.h>
#include <string.h>
#include "sc_core.cuh"

namespace sc {

inline int build_tables(const sc_model *m, Tables *T, char *err, size_t errlen) {
#define SC_FAIL(...) do { snprintf(err, errlen, __VA_ARGS__); return SC_EINVAL; } while (0)
    memset(T, 0, sizeof(*T));
    SelfTables *ST = &T->self;
    if (!m) SC_FAIL("null model");
    if (m->nb < 1 || m->nb > NB) SC_FAIL("nb=%d outside 1..%d", m->nb, NB);
    if (m->nlev < 1 || m->nlev > MAXLEV) SC_FAIL("nlev=%d outside 1..%d", m->nlev, MAXLEV);
    if (m->ncand < 0 || m->ncand > SC_MAX_CAND) SC_FAIL("ncand=%d > %d", m->ncand, SC_MAX_CAND);
    if (m->ncp < 1 || m->ncp > SC_MAX_CPOINTS) SC_FAIL("ncp=%d outside 1..%d", m->ncp, SC_MAX_CPOINTS);
    if (m->nee < 1 || m->nee > 4) SC_FAIL("nee=%d outside 1..4", m->nee);
    if (m->max_contacts < 0 || m->max_contacts > MAXC) SC_FAIL("max_contacts=%d > %d", m->max_contacts, MAXC);
    if (m->num_substeps < 1 || m->num_solver_iters < 0 || !(m->dt > 0)) SC_FAIL("bad engine settings");
    if (7 + 4 * (m->nb - 1) + 6 + 3 * (m->nb - 1) != SC_STATE_DIM) SC_FAIL("state layout needs %d joints", (SC_STATE_DIM - 13) / 7);
    T->nb = m->nb; T->nlev = m->nlev; T->ncand = m->ncand; T->ncp = m->ncp; T->nee = m->nee;
    T->niter = m->num_solver_iters; T->nsub = m->num_substeps; T->maxc = m->max_contacts; T->nj = m->nb - 1;
    for (int d = 0; d <= m->nlev; d++) T->level_start[d] = m->level_start[d];
    if (T->level_start[0] != 0 || T->level_start[1] != 1 || T->level_start[m->nlev] != m->nb) SC_FAIL("bad level_start");
    for (int d = 0; d < m->nlev; d++) {
        int w = T->level_start[d + 1] - T->level_start[d];
        if (w < 1 || w > 4) SC_FAIL("tree level %d has %d bodies (1..4 supported)", d, w);
    }
    for (int b = 0; b < m->nb; b++) {
        int p = m->parent[b];
        if ((b == 0) != (p < 0) || p >= b) SC_FAIL("bodies must be listed parents-first (body %d)", b);
        T->parent[b] = p;
        T->depth[b] = b ? T->depth[p] + 1 : 0;
        if (b < T->level_start[T->depth[b]] || b >= T->level_start[T->depth[b] + 1]) SC_FAIL("body %d not in its level", b);
        T->icslot[b] = (T->depth[b] & 1) * ICW + (b - T->level_start[T->depth[b]]);
        T->anc_mask[b] = (1 << b) | (b ? T->anc_mask[p] : 0);
        for (int d = 0; d < MAXLEV; d++) T->anc[b][d] = d < T->depth[b] ? T->anc[p][d] : (d == T->depth[b] ? b : -1);
        if (b) {
            if (T->nchild[p] >= 4) SC_FAIL("body %d has more than 4 children", p);
            T->child[p][T->nchild[p]++] = b;
            int j = m->jidx[b];
            if (j < 0 || j >= m->nb - 1) SC_FAIL("bad jidx for body %d", b);
            T->jidx[b] = j; T->jbody[j] = b;
        } else T->jidx[b] = -1;
        for (int k = 0; k < 3; k++) T->jorg[3 * b + k] = m->jorg[3 * b + k];
        for (int k = 0; k < 10; k++) { T->idyn[10 * b + k] = m->inertia_dyn[10 * b + k]; T->ipd[10 * b + k] = m->inertia_pd[10 * b + k]; }
        T->kp[b] = m->kp[b]; T->kd[b] = m->kd[b]; T->maxf[b] = m->maxf[b]; T->jw[b] = m->joint_weights[b];
    }
    T->cand_per_lane = (m->ncand + LPS - 1) / LPS;
    for (int c = 0; c < m->ncand; c++) {
        if (m->cand_body[c] < 0 || m->cand_body[c] >= m->nb) SC_FAIL("bad cand_body");
        T->cand_body[c] = m->cand_body[c];
        for (int k = 0; k < 4; k++) T->cand[4 * c + k] = m->cand[4 * c + k];
    }
    double tm = 0;
    for (int i = 0; i < m->ncp; i++) {
        if (m->cp_body[i] < 0 || m->cp_body[i] >= m->nb) SC_FAIL("bad cp_body");
        T->cp_body[i] = m->cp_body[i]; T->cp_mass[i] = m->cp_mass[i]; tm += m->cp_mass[i];
        for (int k = 0; k < 3; k++) T->cp_off[3 * i + k] = m->cp_off[3 * i + k];
    }
    for (int e = 0; e < m->nee; e++) {
        if (m->ee[e] < 0 || m->ee[e] >= m->ncp) SC_FAIL("bad end effector %d", m->ee[e]);
        T->ee[e] = m->ee[e];
    }
    for (int k = 0; k < 5; k++) T->cw[k] = m->cost_weights[k];
    T->height = m->height; T->dt = m->dt; T->h = m->dt / m->num_substeps;
    for (int k = 0; k < 3; k++) T->g[k] = m->gravity[k];
    T->mu = m->friction; T->mu_self = m->friction_self; T->erp = m->erp; T->cthresh = m->contact_threshold; T->maxvel = m->max_velocity;
    T->inv_total_mass = (float)(1.0 / tm);
    if (!(m->warmstart >= 0.0f && m->warmstart <= 1.0f)) SC_FAIL("warmstart=%g outside 0..1", (double)m->warmstart);
    T->warmstart = m->warmstart;
    if (m->spinning_friction != 0.0f || m->spinning_friction_self != 0.0f)
        SC_FAIL("spinning friction is not implemented on the device (3 rows per contact); use 0");
    // self collision: primitives with bounding spheres, pairs with their broad-phase radius
    if (m->npair < 0 || m->npair > NPAIR) SC_FAIL("npair=%d outside 0..%d", m->npair, NPAIR);
    if (m->npair > 0) {
        if (m->nprim < 1 || m->nprim > NPRIM || !m->prim || !m->prim_body || !m->pair) SC_FAIL("nprim=%d outside 1..%d", m->nprim, NPRIM);
        ST->nprim = m->nprim; ST->npair = m->npair;
        ST->prims_per_lane = (m->nprim + LPS - 1) / LPS;
        ST->pairs_per_lane = (m->npair + LPS - 1) / LPS;
        for (int p = 0; p < m->nprim; p++) {
            if (m->prim_body[p] < 0 || m->prim_body[p] >= m->nb) SC_FAIL("bad prim_body");
            ST->prim_body[p] = (unsigned char)m->prim_body[p];
            const auto *q = m->prim + 7 * p;      // (auto: tests/emu/emu64.cpp compiles this file with float := double)
            for (int k = 0; k < 7; k++) ST->prim[8 * p + k] = q[k];
            const double dx = q[3] - q[0], dy = q[4] - q[1], dz = q[5] - q[2];
            ST->prim[8 * p + 7] = (float)(q[6] + 0.5 * sqrt(dx * dx + dy * dy + dz * dz));
        }
        if (m->nbox < 0 || m->nbox > SC_MAX_BOXES || (m->nbox > 0 && (!m->prim_box || !m->box))) SC_FAIL("nbox=%d outside 0..%d", m->nbox, SC_MAX_BOXES);
        ST->nbox = m->nbox;
        for (int p = 0; p < m->nprim; p++) {
            const int bx = m->nbox > 0 ? m->prim_box[p] : -1;
            if (bx >= m->nbox) SC_FAIL("bad prim_box");
            ST->prim_box[p] = (signed char)(bx < 0 ? -1 : bx);
            if (bx >= 0) {
                const auto *q = m->box + 15 * bx;
                for (int k = 0; k < 15; k++) ST->box[15 * bx + k] = q[k];
                // the broad phase's bounding sphere has to hold the box, not just its capsule stand-in
                ST->prim[8 * p + 7] = (float)sqrt((double)q[12] * q[12] + (double)q[13] * q[13] + (double)q[14] * q[14]);
                // the capsule around the long axis that contains the box
                int kx = 0;
                for (int k = 1; k < 3; k++) if (q[12 + k] > q[12 + kx]) kx = k;
                double rc = 0;
                for (int k = 0; k < 3; k++) {
                    const double ax = (double)q[3 + 3 * k + kx] * q[12 + kx];          // column kx of R, scaled by the half length
                    ST->box_core[7 * bx + k] = (float)(q[k] + ax);
                    ST->box_core[7 * bx + 3 + k] = (float)(q[k] - ax);
                    if (k != kx) rc += (double)q[12 + k] * q[12 + k];
                }
                ST->box_core[7 * bx + 6] = (float)(sqrt(rc) * 1.0001);
                // a box whose axes are not its body's carries the flag "rotated" in the sign of this bound
                bool ident = true;
                for (int k = 0; k < 9; k++) if (fabs((double)q[3 + k] - (k % 4 == 0 ? 1.0 : 0.0)) > 1e-7) ident = false;
                if (!ident) ST->box_core[7 * bx + 6] = -ST->box_core[7 * bx + 6];
            }
        }
        for (int i = 0; i < m->npair; i++) {
            const int a = m->pair[2 * i], b = m->pair[2 * i + 1];
            if (a < 0 || b < 0 || a >= m->nprim || b >= m->nprim) SC_FAIL("bad pair %d", i);
            ST->pair_a[i] = (unsigned char)a; ST->pair_b[i] = (unsigned char)b;
            const double r = (double)ST->prim[8 * a + 7] + ST->prim[8 * b + 7] + m->contact_threshold;
            ST->pair_r2[i] = (float)(r * r * 1.0001);
        }
    }
    return SC_OK;
#undef SC_FAIL
}

}  // namespace sc

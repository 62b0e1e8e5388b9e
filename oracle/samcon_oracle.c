/*
 * samcon_oracle.c -- CPU restatement (float64, plain C) of SamCon's per-window hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under samcon_b200/ may import, link or call this file;
 * it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs have an independent answer to compare the CUDA path with.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party wheel
 * pybullet==3.2.5 (/root/reference/requirements.txt:5), which is neither vendored nor
 * installable in the build image, and the reference's result pickles are absent
 * (/root/reference/.MISSING_LARGE_BLOBS:4-9).  What follows restates the reference's own
 * Python (cited per function) and the published algorithms PyBullet implements
 * (Featherstone RNEA/CRBA, Tan et al. 2011 stable PD, Bullet's sequential-impulse / PGS
 * multibody contact rows), with every Bullet-internal choice listed in DESIGN.md.
 *
 * Deliberately written with *different algorithms* than the CUDA kernels: dense 6x6
 * spatial algebra, CRBA + dense Cholesky for every solve, joint-space contact Jacobians,
 * velocity-space PGS.  The kernels use an O(n) articulated-body recursion and a
 * contact-space Gauss-Seidel; agreement between the two is therefore a real check.
 *
 * Conventions: quaternions xyzw; rotation matrices row-major, v_world = R v_link; spatial
 * vectors [angular; linear] in link coordinates about the link-frame origin (= URDF joint
 * origin); generalised velocity = [base twist in base coords (6); spherical joint omegas in
 * child coords (3 each, URDF order)].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 32
#define MAXS 64
#define MAXV 102
#define MAXCAND 256
#define MAXROW (4 * MAXCAND)
#define MAXCC 64   /* contact-cache entries per sample (>= maxc) */
#define MAXPAIR 512

typedef struct {
    int nl, parent[MAXL], jtype[MAXL], dof[MAXL], nv, nj, jlink[MAXL];
    double jorg[MAXL][3], com[MAXL][3], mass[MAXL];
    double I6[2][MAXL][36];
    int ns, sh_link[MAXS], sh_kind[MAXS];
    double sh_dims[MAXS][3], sh_pos[MAXS][3], sh_rot[MAXS][9];
    double kp[MAXL], kd[MAXL], maxf[MAXL];
    double dt; int nsub, niter; double g[3], mu, erp, cthresh; int maxc; double maxvel;
    double jw[MAXL]; int nee, ee[8]; double cw[5], height;
    /* self collision (URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS, humanoid_stable_pd.py:10-13) */
    int selfcol, npair, pair_a[MAXPAIR], pair_b[MAXPAIR];
    double mu_self, cap_p0[MAXS][3], cap_p1[MAXS][3], cap_r[MAXS];
    /* contact persistence: warm-start factor applied to the cached normal impulse of a contact that was also present in
     * the previous sub-step (0 = stateless contacts); spinning (torsional) friction coefficients, ground / link-link */
    double warmstart, spin, spin_self;
    int exact_box;     /* link-vs-link narrow phase: boxes as boxes (1) or as their capsule stand-ins (0, round 1) */
} oc_model;

typedef struct {
    double p[3], Q[4], q[MAXL][4], vw[3], ww[3], w[MAXL][3];
    /* contact cache: candidate id and normal impulse of every contact of the previous sub-step (what Bullet keeps in its
     * persistent manifolds, btManifoldPoint::m_appliedImpulse; saved / restored with the world by saveBullet / restoreState,
     * /root/reference/envs/humanoid_tracking_env.py:178-182) */
    int ncache, cid[MAXCC];
    double clam[MAXCC];
} oc_state;

typedef struct {
    double R[MAXL][9], p[MAXL][3], E[MAXL][9], v[MAXL][6];
} oc_kin;

/* ------------------------------------------------------------------ small algebra */
static void cross(const double *a, const double *b, double *c) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    c[0] = x; c[1] = y; c[2] = z;
}
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void mv3(const double *R, const double *v, double *o) {
    double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
           z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void mtv3(const double *R, const double *v, double *o) {
    double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
           z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
static void mm3(const double *A, const double *B, double *C) {
    double t[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
        t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, t, sizeof t);
}
static void quat_to_R(const double *q, double *R) {
    double x = q[0], y = q[1], z = q[2], w = q[3];
    double n = 1.0 / sqrt(x * x + y * y + z * z + w * w);
    x *= n; y *= n; z *= n; w *= n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static void qmul(const double *a, const double *b, double *o) { /* o = a (x) b, xyzw */
    double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    double y = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    double z = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
static void qnormalize(double *q) {
    double n = 1.0 / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] *= n;
}
/* state rows coming back from a previous call are unit to rounding and are left alone (a resumed window then repeats the
 * uncut one exactly); anything else is normalised.  The threshold is the device's (float32 states). */
static void qnormalize_input(double *q) {
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (fabs(n2 - 1.0) > 1e-6) qnormalize(q);
}
/* exponential-map increment, Bullet's multibody quaternion update with its small-angle Taylor branch
 * [BULLET-UPSTREAM, unverified] */
static void quat_exp(const double *w, double h, double *dq) {
    double a = sqrt(dot3(w, w)), s;
    if (a < 1e-3) s = 0.5 * h - h * h * h * 0.020833333333 * a * a;
    else s = sin(0.5 * a * h) / a;
    dq[0] = w[0] * s; dq[1] = w[1] * s; dq[2] = w[2] * s; dq[3] = cos(0.5 * a * h);
}

/* motion transform parent->child: w' = E w, v' = E (v - r x w) */
static void xmot(const double *E, const double *r, const double *v, double *o) {
    double t[3], u[3];
    cross(r, v, t);
    for (int k = 0; k < 3; k++) u[k] = v[3 + k] - t[k];
    mv3(E, v, o);
    mv3(E, u, o + 3);
}
/* force transform child->parent (transpose of xmot), accumulating */
static void xforce_acc(const double *E, const double *r, const double *f, double *acc) {
    double n[3], l[3], t[3];
    mtv3(E, f, n);
    mtv3(E, f + 3, l);
    cross(r, l, t);
    for (int k = 0; k < 3; k++) { acc[k] += n[k] + t[k]; acc[3 + k] += l[k]; }
}
static void crm(const double *v, const double *m, double *o) {
    double a[3], b[3], c[3];
    cross(v, m, a); cross(v, m + 3, b); cross(v + 3, m, c);
    for (int k = 0; k < 3; k++) { o[k] = a[k]; o[3 + k] = b[k] + c[k]; }
}
static void crf(const double *v, const double *f, double *o) {
    double a[3], b[3], c[3];
    cross(v, f, a); cross(v + 3, f + 3, b); cross(v, f + 3, c);
    for (int k = 0; k < 3; k++) { o[k] = a[k] + b[k]; o[3 + k] = c[k]; }
}
static void mv6(const double *I, const double *v, double *o) {
    double t[6];
    for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += I[6 * i + j] * v[j]; t[i] = s; }
    memcpy(o, t, sizeof t);
}
/* dense 6x6 motion-transform matrix X = [E 0; -E rx E] */
static void xmat(const double *E, const double *r, double *X) {
    double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0}, Erx[9];
    mm3(E, rx, Erx);
    memset(X, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        X[6 * i + j] = E[3 * i + j];
        X[6 * (3 + i) + 3 + j] = E[3 * i + j];
        X[6 * (3 + i) + j] = -Erx[3 * i + j];
    }
}

/* ------------------------------------------------------------------ model */
static void spatial_inertia(double m, const double *c, const double *Ic, double *I6) {
    double cx[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0}, cxcx[9];
    mm3(cx, cx, cxcx);
    memset(I6, 0, 36 * sizeof(double));
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        I6[6 * i + j] = Ic[3 * i + j] - m * cxcx[3 * i + j];
        I6[6 * i + 3 + j] = m * cx[3 * i + j];
        I6[6 * (3 + i) + j] = -m * cx[3 * i + j];
    }
    for (int i = 0; i < 3; i++) I6[6 * (3 + i) + 3 + i] = m;
}

/* Self-collision geometry.  Every collision primitive is stood in for by a sphere-swept segment (p0, p1, r) in its
 * link frame: spheres and capsules exactly; a box by the capsule along its longest side with the same reach along
 * that side and the same cross-section area (r = sqrt(s1 s2) / 2 for the two short sides) -- DESIGN.md §2 lists
 * this as a deviation from Bullet's box-box / box-convex narrow phase.  Pairs: shapes of two different links that
 * are not in an ancestor relation (EXCLUDE_ALL_PARENTS), in shape order. */
static void self_collision_setup(oc_model *m) {
    for (int s = 0; s < m->ns; s++) {
        double e[3] = {0, 0, 0}, a[3], r;
        if (m->sh_kind[s] == 0) r = m->sh_dims[s][0];
        else if (m->sh_kind[s] == 1) { r = m->sh_dims[s][0]; e[2] = 0.5 * m->sh_dims[s][1]; }
        else {
            int k = 0;
            for (int i = 1; i < 3; i++) if (m->sh_dims[s][i] > m->sh_dims[s][k]) k = i;
            double prod = 1;
            for (int i = 0; i < 3; i++) if (i != k) prod *= m->sh_dims[s][i];
            r = 0.5 * sqrt(prod);
            e[k] = 0.5 * m->sh_dims[s][k] - r;
            if (e[k] < 0) e[k] = 0;
        }
        mv3(m->sh_rot[s], e, a);
        for (int c = 0; c < 3; c++) { m->cap_p0[s][c] = m->sh_pos[s][c] + a[c]; m->cap_p1[s][c] = m->sh_pos[s][c] - a[c]; }
        m->cap_r[s] = r;
    }
    m->npair = 0;
    if (!m->selfcol) return;
    for (int a = 0; a < m->ns; a++) for (int b = a + 1; b < m->ns; b++) {
        int la = m->sh_link[a], lb = m->sh_link[b], rel = la == lb;
        for (int p = m->parent[la]; p >= 0 && !rel; p = m->parent[p]) if (p == lb) rel = 1;
        for (int p = m->parent[lb]; p >= 0 && !rel; p = m->parent[p]) if (p == la) rel = 1;
        if (!rel && m->npair < MAXPAIR) { m->pair_a[m->npair] = a; m->pair_b[m->npair] = b; m->npair++; }
    }
}

oc_model *oc_create(int nl, const int *parent, const int *jtype, const double *jorg, const double *com,
                    const double *mass, const double *com_dyn, const double *I_dyn, const double *com_pd,
                    const double *I_pd, int ns, const int *sh_link, const int *sh_kind, const double *sh_dims,
                    const double *sh_pos, const double *sh_rot, const double *kp, const double *kd,
                    const double *maxf,
                    const double *simp /* dt,nsub,niter,gx,gy,gz,mu,erp,cthresh,maxc,maxvel,selfcol,mu_self,warmstart,spin,spin_self,exact_box */,
                    const double *jw, int nee, const int *ee, const double *cw, double height) {
    if (nl > MAXL || ns > MAXS) return NULL;
    oc_model *m = (oc_model *)calloc(1, sizeof(oc_model));
    m->nl = nl; m->ns = ns;
    int nv = 6, nj = 0;
    for (int i = 0; i < nl; i++) {
        m->parent[i] = parent[i]; m->jtype[i] = jtype[i]; m->mass[i] = mass[i];
        memcpy(m->jorg[i], jorg + 3 * i, 24); memcpy(m->com[i], com + 3 * i, 24);
        spatial_inertia(mass[i], com_dyn + 3 * i, I_dyn + 9 * i, m->I6[0][i]);
        spatial_inertia(mass[i], com_pd + 3 * i, I_pd + 9 * i, m->I6[1][i]);
        if (i == 0) m->dof[i] = 0;
        else if (jtype[i] == 1) { m->dof[i] = nv; nv += 3; m->jlink[nj++] = i; }
        else m->dof[i] = -1;
    }
    if (nv > MAXV) { free(m); return NULL; }
    m->nv = nv; m->nj = nj;
    for (int s = 0; s < ns; s++) {
        m->sh_link[s] = sh_link[s]; m->sh_kind[s] = sh_kind[s];
        memcpy(m->sh_dims[s], sh_dims + 3 * s, 24); memcpy(m->sh_pos[s], sh_pos + 3 * s, 24);
        memcpy(m->sh_rot[s], sh_rot + 9 * s, 72);
    }
    for (int k = 0; k < nj; k++) { m->kp[k] = kp[k]; m->kd[k] = kd[k]; m->maxf[k] = maxf[k]; m->jw[k] = jw[k]; }
    m->dt = simp[0]; m->nsub = (int)simp[1]; m->niter = (int)simp[2];
    m->g[0] = simp[3]; m->g[1] = simp[4]; m->g[2] = simp[5];
    m->mu = simp[6]; m->erp = simp[7]; m->cthresh = simp[8]; m->maxc = (int)simp[9]; m->maxvel = simp[10];
    m->nee = nee; for (int k = 0; k < nee; k++) m->ee[k] = ee[k];
    memcpy(m->cw, cw, 40); m->height = height;
    m->selfcol = (int)simp[11]; m->mu_self = simp[12];
    m->warmstart = simp[13]; m->spin = simp[14]; m->spin_self = simp[15]; m->exact_box = (int)simp[16];
    self_collision_setup(m);
    return m;
}
void oc_destroy(oc_model *m) { free(m); }
int oc_nv(const oc_model *m) { return m->nv; }

/* 132-float layout: /root/reference/envs/humanoid_stable_pd.py:137-165 */
static void unpack(const oc_model *m, const double *s, oc_state *st) {
    int nj = m->nj;
    memcpy(st->p, s, 24); memcpy(st->Q, s + 3, 32);
    for (int k = 0; k < nj; k++) memcpy(st->q[k], s + 7 + 4 * k, 32);
    memcpy(st->vw, s + 7 + 4 * nj, 24); memcpy(st->ww, s + 10 + 4 * nj, 24);
    for (int k = 0; k < nj; k++) memcpy(st->w[k], s + 13 + 4 * nj + 3 * k, 24);
    st->ncache = 0;
}
/* contact-cache row (2 * maxc values): candidate ids (-1 = empty slot), then their normal impulses */
static void cache_load(const oc_model *m, const double *row, oc_state *st) {
    st->ncache = 0;
    if (!row) return;
    for (int i = 0; i < m->maxc && i < MAXCC; i++)
        if (row[i] >= 0) { st->cid[st->ncache] = (int)row[i]; st->clam[st->ncache] = row[m->maxc + i]; st->ncache++; }
}
static void cache_store(const oc_model *m, const oc_state *st, double *row) {
    if (!row) return;
    for (int i = 0; i < m->maxc; i++) {
        row[i] = i < st->ncache ? st->cid[i] : -1.0;
        row[m->maxc + i] = i < st->ncache ? st->clam[i] : 0.0;
    }
}
static void pack(const oc_model *m, const oc_state *st, double *s) {
    int nj = m->nj;
    memcpy(s, st->p, 24); memcpy(s + 3, st->Q, 32);
    for (int k = 0; k < nj; k++) memcpy(s + 7 + 4 * k, st->q[k], 32);
    memcpy(s + 7 + 4 * nj, st->vw, 24); memcpy(s + 10 + 4 * nj, st->ww, 24);
    for (int k = 0; k < nj; k++) memcpy(s + 13 + 4 * nj + 3 * k, st->w[k], 24);
}

/* ------------------------------------------------------------------ kinematics */
static void fk(const oc_model *m, const oc_state *st, oc_kin *k) {
    quat_to_R(st->Q, k->R[0]);
    memcpy(k->p[0], st->p, 24);
    for (int i = 0; i < 9; i++) k->E[0][i] = k->R[0][(i % 3) * 3 + i / 3];
    mtv3(k->R[0], st->ww, k->v[0]);
    mtv3(k->R[0], st->vw, k->v[0] + 3);
    int j = 0;
    for (int i = 1; i < m->nl; i++) {
        int p = m->parent[i];
        double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3], wj[3] = {0, 0, 0};
        if (m->jtype[i] == 1) { quat_to_R(st->q[j], Rj); memcpy(wj, st->w[j], 24); j++; }
        mm3(k->R[p], Rj, k->R[i]);
        mv3(k->R[p], m->jorg[i], t);
        for (int c = 0; c < 3; c++) k->p[i][c] = k->p[p][c] + t[c];
        for (int c = 0; c < 9; c++) k->E[i][c] = Rj[(c % 3) * 3 + c / 3];
        xmot(k->E[i], m->jorg[i], k->v[p], k->v[i]);
        for (int c = 0; c < 3; c++) k->v[i][c] += wj[c];
    }
}
static void gen_vel(const oc_model *m, const oc_state *st, const oc_kin *k, double *v) {
    memcpy(v, k->v[0], 48);
    for (int j = 0; j < m->nj; j++) memcpy(v + 6 + 3 * j, st->w[j], 24);
}

/* RNEA: tau = M qdd + C(q, v) (+ gravity), inertia set `set` */
static void rnea(const oc_model *m, int set, const oc_state *st, const oc_kin *k, const double *qdd, int grav,
                 double *tau) {
    double a[MAXL][6], f[MAXL][6], t[6], u[6];
    double ag[3] = {0, 0, 0};
    if (grav) for (int c = 0; c < 3; c++) ag[c] = -m->g[c];
    memset(a[0], 0, 24);
    mtv3(k->R[0], ag, a[0] + 3);
    if (qdd) for (int c = 0; c < 6; c++) a[0][c] += qdd[c];
    int j = 0;
    for (int i = 0; i < m->nl; i++) {
        if (i > 0) {
            xmot(k->E[i], m->jorg[i], a[m->parent[i]], a[i]);
            if (m->jtype[i] == 1) {
                double vj[6] = {st->w[j][0], st->w[j][1], st->w[j][2], 0, 0, 0};
                crm(k->v[i], vj, t);
                for (int c = 0; c < 6; c++) a[i][c] += t[c];
                if (qdd) for (int c = 0; c < 3; c++) a[i][c] += qdd[m->dof[i] + c];
                j++;
            }
        }
        mv6(m->I6[set][i], a[i], f[i]);
        mv6(m->I6[set][i], k->v[i], t);
        crf(k->v[i], t, u);
        for (int c = 0; c < 6; c++) f[i][c] += u[c];
    }
    for (int i = m->nl - 1; i > 0; i--) {
        if (m->jtype[i] == 1) for (int c = 0; c < 3; c++) tau[m->dof[i] + c] = f[i][c];
        xforce_acc(k->E[i], m->jorg[i], f[i], f[m->parent[i]]);
    }
    memcpy(tau, f[0], 48);
}

/* CRBA: joint-space mass matrix (nv x nv, row-major, full) */
static void crba(const oc_model *m, int set, const oc_kin *k, double *M) {
    static __thread double Ic[MAXL][36];
    int nv = m->nv;
    memcpy(Ic, m->I6[set], sizeof(double) * 36 * m->nl);
    for (int i = m->nl - 1; i > 0; i--) {
        double X[36], T[36];
        xmat(k->E[i], m->jorg[i], X);
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
            double s = 0; for (int e = 0; e < 6; e++) s += Ic[i][6 * r + e] * X[6 * e + c];
            T[6 * r + c] = s;
        }
        double *P = Ic[m->parent[i]];
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
            double s = 0; for (int e = 0; e < 6; e++) s += X[6 * e + r] * T[6 * e + c];
            P[6 * r + c] += s;
        }
    }
    memset(M, 0, sizeof(double) * nv * nv);
    for (int i = 0; i < m->nl; i++) {
        if (m->dof[i] < 0) continue;
        int ni = i == 0 ? 6 : 3, di = m->dof[i];
        double F[6][6];
        for (int c = 0; c < ni; c++) for (int r = 0; r < 6; r++) F[c][r] = Ic[i][6 * r + c];
        for (int c = 0; c < ni; c++) for (int r = 0; r < ni; r++) M[(di + r) * nv + di + c] = F[c][r];
        int jl = i;
        while (m->parent[jl] >= 0) {
            for (int c = 0; c < ni; c++) {
                double acc[6] = {0, 0, 0, 0, 0, 0};
                xforce_acc(k->E[jl], m->jorg[jl], F[c], acc);
                memcpy(F[c], acc, 48);
            }
            jl = m->parent[jl];
            if (m->dof[jl] < 0) continue;
            int nj = jl == 0 ? 6 : 3, dj = m->dof[jl];
            for (int c = 0; c < ni; c++) for (int r = 0; r < nj; r++) {
                M[(di + c) * nv + dj + r] = F[c][r];
                M[(dj + r) * nv + di + c] = F[c][r];
            }
        }
    }
}

static int cholesky(double *A, int n) { /* in place, lower */
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (d <= 0) return -1;
        d = sqrt(d); A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    return 0;
}
static void chol_solve(const double *L, int n, const double *b, double *x) {
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= L[i * n + k] * x[k];
        x[i] = s / L[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k * n + i] * x[k];
        x[i] = s / L[i * n + i];
    }
}

/* ------------------------------------------------------------------ stable PD
 * /root/reference/envs/humanoid_stable_pd.py:124-135 -> STABLE_PD_CONTROL.  Restates SURVEY.md §9.3:
 * acc = (M + dt Kd)^-1 (Kp err - Kd v - C), tau = Kp err - Kd (v + dt acc), clamp +-maxForce;
 * err = axis-angle of (q (+) dt qdot)^-1 (x) q_target in the predicted child frame; root rows have no gains. */
static void pd_error(const double *q, const double *w, const double *qt, double dt, double *err) {
    double wq[4] = {w[0], w[1], w[2], 0}, dq[4], qp[4], qc[4], d[4];
    qmul(q, wq, dq);
    for (int c = 0; c < 4; c++) qp[c] = q[c] + 0.5 * dt * dq[c];
    qnormalize(qp);
    qc[0] = -qp[0]; qc[1] = -qp[1]; qc[2] = -qp[2]; qc[3] = qp[3];
    double tn[4] = {qt[0], qt[1], qt[2], qt[3]};
    qnormalize(tn);
    qmul(qc, tn, d);
    if (d[3] < 0) for (int c = 0; c < 4; c++) d[c] = -d[c];
    double s = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (s < 1e-12) { err[0] = err[1] = err[2] = 0; return; }
    double ang = 2.0 * atan2(s, d[3]);
    for (int c = 0; c < 3; c++) err[c] = d[c] / s * ang;
}
static void stable_pd(const oc_model *m, const oc_state *st, const oc_kin *k, const double *action, double *tau) {
    static __thread double M[MAXV * MAXV];
    double C[MAXV], rhs[MAXV], acc[MAXV], pterm[MAXV];
    int nv = m->nv;
    crba(m, 1, k, M);
    rnea(m, 1, st, k, NULL, 1, C);
    for (int c = 0; c < 6; c++) { rhs[c] = -C[c]; pterm[c] = 0; }
    for (int j = 0; j < m->nj; j++) {
        double err[3];
        int d = 6 + 3 * j;
        pd_error(st->q[j], st->w[j], action + 4 * j, m->dt, err);
        for (int c = 0; c < 3; c++) {
            pterm[d + c] = m->kp[j] * err[c] - m->kd[j] * st->w[j][c];
            rhs[d + c] = pterm[d + c] - C[d + c];
            M[(d + c) * nv + d + c] += m->dt * m->kd[j];
        }
    }
    cholesky(M, nv);
    chol_solve(M, nv, rhs, acc);
    for (int c = 0; c < 6; c++) tau[c] = 0;
    for (int j = 0; j < m->nj; j++) for (int c = 0; c < 3; c++) {
        int d = 6 + 3 * j + c;
        double t = pterm[d] - m->dt * m->kd[j] * acc[d];
        if (t > m->maxf[j]) t = m->maxf[j];
        if (t < -m->maxf[j]) t = -m->maxf[j];
        tau[d] = t;
    }
}

/* ------------------------------------------------------------------ ground contact
 * plane_implicit.urdf at z = 0 (/root/reference/envs/humanoid_tracking_env.py:75-76).  Stateless
 * manifold: sphere -> lowest point, capsule -> lowest point of each cap, box -> corners; a point joins
 * the solve when its height is below the contact threshold; deepest `maxc` kept (ties: lower index). */
typedef struct { int link, link2, id; double pos[3], pos2[3], n[3], dist, mu, spin; } oc_contact;

static double clamp01(double x) { return x < 0 ? 0 : (x > 1 ? 1 : x); }
/* closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9); parallel segments: s = 0 */
static void seg_seg(const double *p1, const double *q1, const double *p2, const double *q2, double *c1, double *c2) {
    double d1[3], d2[3], r[3], s, t;
    for (int k = 0; k < 3; k++) { d1[k] = q1[k] - p1[k]; d2[k] = q2[k] - p2[k]; r[k] = p1[k] - p2[k]; }
    double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
    const double EPS = 1e-12;
    if (a <= EPS && e <= EPS) s = t = 0;
    else if (a <= EPS) { s = 0; t = clamp01(f / e); }
    else {
        double c = dot3(d1, r);
        if (e <= EPS) { t = 0; s = clamp01(-c / a); }
        else {
            double b = dot3(d1, d2), den = a * e - b * b;
            s = den > 1e-6 * a * e ? clamp01((b * f - c * e) / den) : 0;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = clamp01(-c / a); }
            else if (t > 1) { t = 1; s = clamp01((b - c) / a); }
        }
    }
    for (int k = 0; k < 3; k++) { c1[k] = p1[k] + d1[k] * s; c2[k] = p2[k] + d2[k] * t; }
}

/* Closest points of a box (half extents h, its own frame) and a segment p0 + t (p1 - p0), t in [0,1] (a sphere centre is
 * the segment with p1 = p0).  f(t) = squared distance of the segment point to the box is convex and piecewise quadratic:
 * safeguarded Newton on f' (each step solves the current piece exactly; a step that leaves the bracket is replaced by a
 * bisection), a fixed number of iterations.  Returns the distance between the cores; q = point of the box, p = point of the
 * segment.  When the segment passes through the box (distance 0) p is the middle of the part inside, q its projection onto
 * the nearest face, *axis / *sgn name that face and the (negative) return value is minus the depth of p below it. */
static double box_segment(const double *h, const double *p0, const double *p1, double *q, double *p, int *axis, double *sgn) {
    double d[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    double dd = dot3(d, d), t = dd > 1e-18 ? clamp01(-dot3(p0, d) / dd) : 0.0, lo = 0.0, hi = 1.0;
    int lo_seen = 0, hi_seen = 0;          /* f' evaluated at the bracket's ends (negative at lo, positive at hi)? */
    for (int it = 0; it < 16; it++) {
        double g1 = 0, g2 = 0;
        for (int i = 0; i < 3; i++) {
            double x = p0[i] + t * d[i], o = x - fmax(-h[i], fmin(h[i], x));
            g1 += o * d[i];
            if (o != 0.0) g2 += d[i] * d[i];
        }
        if (g1 == 0 || g2 <= 0) break;
        if (g1 > 0) { if (t <= 0.0) break; hi = t; hi_seen = 1; }          /* rising at the start of the segment: done */
        else { if (t >= 1.0) break; lo = t; lo_seen = 1; }
        double tn = t - g1 / g2;
        if (tn <= lo) tn = lo_seen ? 0.5 * (lo + hi) : lo;
        else if (tn >= hi) tn = hi_seen ? 0.5 * (lo + hi) : hi;
        if (fabs(tn - t) < 1e-15) { t = tn; break; }
        t = tn;
    }
    double len2 = 0;
    for (int i = 0; i < 3; i++) { p[i] = p0[i] + t * d[i]; q[i] = fmax(-h[i], fmin(h[i], p[i])); len2 += (p[i] - q[i]) * (p[i] - q[i]); }
    *axis = -1; *sgn = 0;
    if (len2 > 1e-24) return sqrt(len2);
    /* the segment enters the box: clip it against the three slabs, take the middle of what is inside */
    double t0 = 0, t1 = 1;
    for (int i = 0; i < 3; i++) {
        if (fabs(d[i]) < 1e-18) continue;
        double a = (-h[i] - p0[i]) / d[i], b = (h[i] - p0[i]) / d[i];
        if (a > b) { double c = a; a = b; b = c; }
        if (a > t0) t0 = a;
        if (b < t1) t1 = b;
    }
    t = t1 >= t0 ? 0.5 * (t0 + t1) : t;
    double best = 1e300;
    for (int i = 0; i < 3; i++) {
        p[i] = p0[i] + t * d[i]; q[i] = p[i];
        double e = h[i] - fabs(p[i]);
        if (e < best) { best = e; *axis = i; }
    }
    *sgn = p[*axis] >= 0 ? 1.0 : -1.0;
    q[*axis] = *sgn * h[*axis];
    return -best;
}

/* narrow phase of a pair in which shape sb is a box and the other side is the sphere-swept segment (a0, a1, ra) in world
 * coordinates: surface points pbox / pseg (world), unit normal from the box to the segment side, signed distance */
static int box_pair(const oc_model *m, const oc_kin *k, int sb, const double *a0, const double *a1, double ra,
                    double *pbox, double *pseg, double *n_box_to_seg, double *dist) {
    int lb = m->sh_link[sb];
    double Rw[9], c[3], t[3], h[3], l0[3], l1[3], q[3], p[3], sgn;
    int axis;
    mm3(k->R[lb], m->sh_rot[sb], Rw);
    mv3(k->R[lb], m->sh_pos[sb], t);
    for (int i = 0; i < 3; i++) { c[i] = k->p[lb][i] + t[i]; h[i] = 0.5 * m->sh_dims[sb][i]; }
    for (int i = 0; i < 3; i++) t[i] = a0[i] - c[i];
    mtv3(Rw, t, l0);
    for (int i = 0; i < 3; i++) t[i] = a1[i] - c[i];
    mtv3(Rw, t, l1);
    double len = box_segment(h, l0, l1, q, p, &axis, &sgn), nl[3];
    if (axis >= 0) { nl[0] = nl[1] = nl[2] = 0; nl[axis] = sgn; }
    else { if (len < 1e-9) return 0; for (int i = 0; i < 3; i++) nl[i] = (p[i] - q[i]) / len; }
    *dist = len - ra;
    double qs[3];
    for (int i = 0; i < 3; i++) qs[i] = p[i] - ra * nl[i];                      /* point of the swept segment's surface */
    mv3(Rw, nl, n_box_to_seg);
    mv3(Rw, q, t); for (int i = 0; i < 3; i++) pbox[i] = c[i] + t[i];
    mv3(Rw, qs, t); for (int i = 0; i < 3; i++) pseg[i] = c[i] + t[i];
    return 1;
}

static int collide(const oc_model *m, const oc_kin *k, oc_contact *out) {
    static __thread oc_contact cand[MAXCAND];
    int n = 0;
    /* link vs link first (candidate order = contact order = Gauss-Seidel row order: pairs, then ground points):
     * one point per primitive pair, normal from the second link to the first */
    for (int pi = 0; pi < m->npair; pi++) {
        int sa = m->pair_a[pi], sb = m->pair_b[pi], la = m->sh_link[sa], lb = m->sh_link[sb];
        double a0[3], a1[3], b0[3], b1[3], c1[3], c2[3], t[3], d[3];
        mv3(k->R[la], m->cap_p0[sa], t); for (int a = 0; a < 3; a++) a0[a] = k->p[la][a] + t[a];
        mv3(k->R[la], m->cap_p1[sa], t); for (int a = 0; a < 3; a++) a1[a] = k->p[la][a] + t[a];
        mv3(k->R[lb], m->cap_p0[sb], t); for (int a = 0; a < 3; a++) b0[a] = k->p[lb][a] + t[a];
        mv3(k->R[lb], m->cap_p1[sb], t); for (int a = 0; a < 3; a++) b1[a] = k->p[lb][a] + t[a];
        if (m->exact_box && (m->sh_kind[sa] == 2 || m->sh_kind[sb] == 2)) {
            /* a box is a box; when both are, the second one is stood in for by its capsule */
            double pbx[3], psg[3], nb2s[3], dist;
            int a_is_box = m->sh_kind[sa] == 2;
            int ok = a_is_box ? box_pair(m, k, sa, b0, b1, m->cap_r[sb], pbx, psg, nb2s, &dist)
                              : box_pair(m, k, sb, a0, a1, m->cap_r[sa], pbx, psg, nb2s, &dist);
            if (ok && dist < m->cthresh && n < MAXCAND) {
                oc_contact *cc = &cand[n++];
                cc->link = la; cc->link2 = lb; cc->dist = dist; cc->mu = m->mu_self; cc->spin = m->spin_self; cc->id = pi;
                for (int a = 0; a < 3; a++) {       /* normal from the second link to the first */
                    cc->n[a] = a_is_box ? -nb2s[a] : nb2s[a];
                    cc->pos[a] = a_is_box ? pbx[a] : psg[a];
                    cc->pos2[a] = a_is_box ? psg[a] : pbx[a];
                }
            }
            continue;
        }
        seg_seg(a0, a1, b0, b1, c1, c2);
        for (int a = 0; a < 3; a++) d[a] = c1[a] - c2[a];
        double len = sqrt(dot3(d, d)), dist = len - m->cap_r[sa] - m->cap_r[sb];
        if (dist < m->cthresh && len > 1e-9 && n < MAXCAND) {
            oc_contact *cc = &cand[n++];
            cc->link = la; cc->link2 = lb; cc->dist = dist; cc->mu = m->mu_self; cc->spin = m->spin_self; cc->id = pi;
            for (int a = 0; a < 3; a++) {
                cc->n[a] = d[a] / len;
                cc->pos[a] = c1[a] - m->cap_r[sa] * cc->n[a];
                cc->pos2[a] = c2[a] + m->cap_r[sb] * cc->n[a];
            }
        }
    }
    /* ground; candidate id = npair + running index of the (shape, point) */
    int gid = m->npair;
    for (int s = 0; s < m->ns; s++) {
        int l = m->sh_link[s];
        double c[3], t[3];
        mv3(k->R[l], m->sh_pos[s], t);
        for (int a = 0; a < 3; a++) c[a] = k->p[l][a] + t[a];
        double Rs[9];
        mm3(k->R[l], m->sh_rot[s], Rs);
        int npt = m->sh_kind[s] == 0 ? 1 : (m->sh_kind[s] == 1 ? 2 : 8);
        for (int q = 0; q < npt; q++) {
            double loc[3] = {0, 0, 0}, r = 0, pw[3];
            if (m->sh_kind[s] == 0) r = m->sh_dims[s][0];
            else if (m->sh_kind[s] == 1) { r = m->sh_dims[s][0]; loc[2] = (q == 0 ? 0.5 : -0.5) * m->sh_dims[s][1]; }
            else {
                loc[0] = ((q >> 2) & 1 ? 0.5 : -0.5) * m->sh_dims[s][0];
                loc[1] = ((q >> 1) & 1 ? 0.5 : -0.5) * m->sh_dims[s][1];
                loc[2] = (q & 1 ? 0.5 : -0.5) * m->sh_dims[s][2];
            }
            mv3(Rs, loc, pw);
            for (int a = 0; a < 3; a++) pw[a] += c[a];
            pw[2] -= r;
            const int this_id = gid++;
            if (pw[2] < m->cthresh && n < MAXCAND) {
                oc_contact *cc = &cand[n++];
                cc->id = this_id; cc->spin = m->spin;
                cc->link = l; cc->link2 = -1; memcpy(cc->pos, pw, 24); memcpy(cc->pos2, pw, 24);
                cc->n[0] = 0; cc->n[1] = 0; cc->n[2] = 1; cc->dist = pw[2]; cc->mu = m->mu;
            }
        }
    }
    if (n <= m->maxc) { memcpy(out, cand, n * sizeof(oc_contact)); return n; }
    /* keep the maxc deepest, preserving candidate order among the kept */
    char keep[MAXCAND] = {0};
    for (int r = 0; r < m->maxc; r++) {
        int best = -1;
        for (int i = 0; i < n; i++) if (!keep[i] && (best < 0 || cand[i].dist < cand[best].dist)) best = i;
        keep[best] = 1;
    }
    int o = 0;
    for (int i = 0; i < n; i++) if (keep[i]) out[o++] = cand[i];
    return o;
}

/* row of the contact Jacobian: d . v_point = J . v_gen */
static void point_row_acc(const oc_model *m, const oc_kin *k, int link, const double *pos, const double *d, double sgn,
                          double *J) {
    for (int a = link; a >= 0; a = m->parent[a]) {
        if (m->dof[a] < 0) continue;
        double rel[3], rxd[3];
        for (int e = 0; e < 3; e++) rel[e] = pos[e] - k->p[a][e];
        cross(rel, d, rxd);
        for (int e = 0; e < 3; e++) {
            double ax[3] = {k->R[a][e], k->R[a][3 + e], k->R[a][6 + e]};
            J[m->dof[a] + e] += sgn * dot3(ax, rxd);
            if (a == 0) J[3 + e] += sgn * dot3(ax, d);
        }
    }
}
/* d . (v_point(link) - v_point2(link2)) = J . v_gen; link2 < 0 is the fixed ground */
static void contact_row(const oc_model *m, const oc_kin *k, const oc_contact *c, const double *d, double *J) {
    memset(J, 0, sizeof(double) * m->nv);
    point_row_acc(m, k, c->link, c->pos, d, 1.0, J);
    if (c->link2 >= 0) point_row_acc(m, k, c->link2, c->pos2, d, -1.0, J);
}
/* spinning-friction row: n . (omega(link) - omega(link2)) = J . v_gen (angular only) */
static void spin_row(const oc_model *m, const oc_kin *k, const oc_contact *c, double *J) {
    memset(J, 0, sizeof(double) * m->nv);
    for (int side = 0; side < 2; side++) {
        int link = side ? c->link2 : c->link;
        double sgn = side ? -1.0 : 1.0;
        if (link < 0) continue;
        for (int a = link; a >= 0; a = m->parent[a]) {
            if (m->dof[a] < 0) continue;
            for (int e = 0; e < 3; e++) {
                double ax[3] = {k->R[a][e], k->R[a][3 + e], k->R[a][6 + e]};
                J[m->dof[a] + e] += sgn * dot3(ax, c->n);
            }
        }
    }
}
/* Bullet's btPlaneSpace1: two unit tangents p, q for a unit normal n */
static void plane_space(const double *n, double *p, double *q) {
    if (fabs(n[2]) > 0.7071067811865475244008443621048490) {
        double a = n[1] * n[1] + n[2] * n[2], k = 1.0 / sqrt(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        double a = n[0] * n[0] + n[1] * n[1], k = 1.0 / sqrt(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

/* ------------------------------------------------------------------ one stepSimulation()
 * /root/reference/envs/humanoid_tracking_env.py:187-189: actuate(action); stepSimulation().
 * numSubSteps internal steps of dt/numSubSteps; torque from the stable-PD solve is held across them. */
static int g_last_contacts = 0;
static double g_last_normal_impulse = 0, g_last_tangent_impulse[2] = {0, 0};   /* diagnostics: last sub-step, ground contacts */

static void substep(const oc_model *m, oc_state *st, const oc_kin *k, const double *tau, double h) {
    static __thread double M[MAXV * MAXV], MinvJt[MAXROW][MAXV], J[MAXROW][MAXV];
    static __thread oc_contact con[MAXCAND];
    double C[MAXV], rhs[MAXV], qdd[MAXV], v[MAXV], dv[MAXV];
    static __thread double b[MAXROW], diag[MAXROW], lam[MAXROW];
    int nv = m->nv;
    int nc = collide(m, k, con);
    g_last_contacts = nc;
    crba(m, 0, k, M);
    rnea(m, 0, st, k, NULL, 1, C);
    for (int i = 0; i < nv; i++) rhs[i] = tau[i] - C[i];
    cholesky(M, nv);
    chol_solve(M, nv, rhs, qdd);
    gen_vel(m, st, k, v);
    /* classical (world-integrated) base linear acceleration: + w x v */
    double wxv[3];
    cross(v, v + 3, wxv);
    for (int c = 0; c < 3; c++) qdd[3 + c] += wxv[c];
    for (int i = 0; i < nv; i++) v[i] += h * qdd[i];

    if (nc > 0) {
        /* rows 3c+d: normal (d = 0) and the two lateral friction directions of contact c; row 3 nc + c: spinning
         * (torsional) friction about the normal, angular only, for contacts with a spinning coefficient */
        int has_spin[MAXCAND], nspin = 0;
        for (int c = 0; c < nc; c++) for (int d = 0; d < 3; d++) {
            int r = 3 * c + d;
            double dirs[3][3]; /* n, btPlaneSpace1(n): (0,0,1) -> (0,-1,0), (1,0,0) */
            memcpy(dirs[0], con[c].n, 24);
            plane_space(con[c].n, dirs[1], dirs[2]);
            contact_row(m, k, &con[c], dirs[d], J[r]);
            chol_solve(M, nv, J[r], MinvJt[r]);
            double dg = 0, rel = 0;
            for (int i = 0; i < nv; i++) { dg += J[r][i] * MinvJt[r][i]; rel += J[r][i] * v[i]; }
            diag[r] = dg;
            double err = -rel;
            if (d == 0) err += con[c].dist > 0 ? -con[c].dist / h : -con[c].dist * m->erp / h;
            b[r] = err / dg;
            lam[r] = 0;
        }
        for (int c = 0; c < nc; c++) {
            int r = 3 * nc + c;
            has_spin[c] = con[c].spin > 0;
            lam[r] = 0;
            if (!has_spin[c]) continue;
            nspin++;
            spin_row(m, k, &con[c], J[r]);
            chol_solve(M, nv, J[r], MinvJt[r]);
            double dg = 0, rel = 0;
            for (int i = 0; i < nv; i++) { dg += J[r][i] * MinvJt[r][i]; rel += J[r][i] * v[i]; }
            diag[r] = dg; b[r] = -rel / dg;
        }
        memset(dv, 0, sizeof(double) * nv);
        /* warm start: a contact that was in the previous sub-step's list starts from `warmstart` x its normal impulse;
         * friction rows start from zero (Bullet's multibody contact setup: isFriction ? 0 : m_appliedImpulse * factor) */
        if (m->warmstart > 0)
            for (int c = 0; c < nc; c++)
                for (int o = 0; o < st->ncache; o++) if (st->cid[o] == con[c].id) {
                    lam[3 * c] = m->warmstart * st->clam[o];
                    for (int i = 0; i < nv; i++) dv[i] += MinvJt[3 * c][i] * lam[3 * c];
                    break;
                }
        for (int it = 0; it < m->niter; it++) {
            for (int pass = 0; pass < 3; pass++)     /* all normals, then the spinning rows, then the friction pairs */
                for (int c = 0; c < nc; c++) for (int d = (pass == 2 ? 1 : 0); d < (pass == 2 ? 3 : 1); d++) {
                    int r = pass == 1 ? 3 * nc + c : 3 * c + d;
                    if (pass == 1 && !has_spin[c]) continue;
                    double jv = 0;
                    for (int i = 0; i < nv; i++) jv += J[r][i] * dv[i];
                    double nl = lam[r] + b[r] - jv / diag[r];
                    double lo = 0, hi = 1e30;
                    if (pass == 2) { hi = con[c].mu * lam[3 * c]; lo = -hi; }
                    if (pass == 1) { hi = con[c].spin * lam[3 * c]; lo = -hi; }
                    if (nl < lo) nl = lo;
                    if (nl > hi) nl = hi;
                    double dl = nl - lam[r];
                    lam[r] = nl;
                    for (int i = 0; i < nv; i++) dv[i] += MinvJt[r][i] * dl;
                }
        }
        (void)nspin;
        for (int i = 0; i < nv; i++) v[i] += dv[i];
        g_last_normal_impulse = 0; g_last_tangent_impulse[0] = g_last_tangent_impulse[1] = 0;
        for (int c = 0; c < nc; c++) if (con[c].link2 < 0) {
            g_last_normal_impulse += lam[3 * c];
            g_last_tangent_impulse[0] += lam[3 * c + 2]; g_last_tangent_impulse[1] -= lam[3 * c + 1];   /* world x, y */
        }
    } else { g_last_normal_impulse = 0; g_last_tangent_impulse[0] = g_last_tangent_impulse[1] = 0; }
    /* the cache now describes this sub-step */
    st->ncache = nc < MAXCC ? nc : MAXCC;
    for (int c = 0; c < st->ncache; c++) { st->cid[c] = con[c].id; st->clam[c] = lam[3 * c]; }
    /* back to the stored representation, clamp, integrate (semi-implicit Euler) */
    mv3(k->R[0], v, st->ww);
    mv3(k->R[0], v + 3, st->vw);
    for (int c = 0; c < 3; c++) {
        st->ww[c] = fmax(-m->maxvel, fmin(m->maxvel, st->ww[c]));
        st->vw[c] = fmax(-m->maxvel, fmin(m->maxvel, st->vw[c]));
    }
    for (int j = 0; j < m->nj; j++) for (int c = 0; c < 3; c++)
        st->w[j][c] = fmax(-m->maxvel, fmin(m->maxvel, v[6 + 3 * j + c]));
    double dq[4], t[4];
    for (int c = 0; c < 3; c++) st->p[c] += h * st->vw[c];
    quat_exp(st->ww, h, dq);
    qmul(dq, st->Q, t); memcpy(st->Q, t, 32); qnormalize(st->Q);
    for (int j = 0; j < m->nj; j++) {
        quat_exp(st->w[j], h, dq);
        qmul(st->q[j], dq, t); memcpy(st->q[j], t, 32); qnormalize(st->q[j]);
    }
}

static void step_state(const oc_model *m, oc_state *st, const double *action, int nsteps) {
    oc_kin k;
    double tau[MAXV];
    double h = m->dt / m->nsub;
    for (int s = 0; s < nsteps; s++) {
        fk(m, st, &k);
        stable_pd(m, st, &k, action, tau);
        for (int u = 0; u < m->nsub; u++) {
            if (u) fk(m, st, &k);
            substep(m, st, &k, tau, h);
        }
    }
}

/* cache: contact-cache row (2 * maxc values) read before and rewritten after the steps, or NULL (empty cache in) */
void oc_step_cache(const oc_model *m, double *state, const double *action, int nsteps, double *cache) {
    oc_state st;
    unpack(m, state, &st);
    cache_load(m, cache, &st);
    qnormalize_input(st.Q);
    for (int j = 0; j < m->nj; j++) qnormalize_input(st.q[j]);
    step_state(m, &st, action, nsteps);
    pack(m, &st, state);
    cache_store(m, &st, cache);
}
void oc_step(const oc_model *m, double *state, const double *action, int nsteps) { oc_step_cache(m, state, action, nsteps, NULL); }
int oc_last_contacts(void) { return g_last_contacts; }
void oc_last_ground_impulse(double *out3) { out3[0] = g_last_tangent_impulse[0]; out3[1] = g_last_tangent_impulse[1]; out3[2] = g_last_normal_impulse; }

/* ------------------------------------------------------------------ cost
 * /root/reference/envs/humanoid_tracking_env.py:78-176 (formulas in SURVEY.md §9.4);
 * CoM: /root/reference/envs/humanoid_stable_pd.py:167-187. */
static void link_coms(const oc_model *m, const oc_state *st, double pos[][3], double vel[][3], double *cpos,
                      double *cvel) {
    oc_kin k;
    fk(m, st, &k);
    double tm = 0;
    for (int c = 0; c < 3; c++) cpos[c] = cvel[c] = 0;
    for (int i = 0; i < m->nl; i++) {
        double t[3], u[3];
        mv3(k.R[i], m->com[i], t);
        for (int c = 0; c < 3; c++) pos[i][c] = k.p[i][c] + t[c];
        cross(k.v[i], m->com[i], u);
        for (int c = 0; c < 3; c++) u[c] += k.v[i][3 + c];
        mv3(k.R[i], u, vel[i]);
        tm += m->mass[i];
        for (int c = 0; c < 3; c++) { cpos[c] += m->mass[i] * pos[i][c]; cvel[c] += m->mass[i] * vel[i][c]; }
    }
    for (int c = 0; c < 3; c++) { cpos[c] /= tm; cvel[c] /= tm; }
}
static double quat_angle(const double *a, const double *b) {
    double na = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
    double nb = sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2] + b[3] * b[3]);
    double d = fabs(a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]) / (na * nb);
    if (d > 1) d = 1;
    return 2.0 * acos(d);
}
void oc_cost(const oc_model *m, const double *sim, const double *kin, double *out /* total,pose,root,ee,bal,com */) {
    oc_state a, b;
    unpack(m, sim, &a); unpack(m, kin, &b);
    double pa[MAXL][3], va[MAXL][3], ca[3], cva[3], pb[MAXL][3], vb[MAXL][3], cb[3], cvb[3];
    link_coms(m, &a, pa, va, ca, cva);
    link_coms(m, &b, pb, vb, cb, cvb);
    double pose = 0;
    for (int j = 0; j < m->nj; j++) {
        double ang = quat_angle(a.q[j], b.q[j]), dw = 0;
        for (int c = 0; c < 3; c++) dw += (a.w[j][c] - b.w[j][c]) * (a.w[j][c] - b.w[j][c]);
        pose += m->jw[j] * (ang * ang + 0.1 * dw);
    }
    pose /= m->nj;
    double ang = quat_angle(a.Q, b.Q), dw = 0;
    for (int c = 0; c < 3; c++) dw += (a.ww[c] - b.ww[c]) * (a.ww[c] - b.ww[c]);
    double root = ang * ang + 0.1 * dw;
    double ee = 0, bal = 0;
    for (int e = 0; e < m->nee; e++) {
        int l = m->ee[e] + 1;
        double dz = pa[l][2] - pb[l][2];
        ee += dz * dz;
        for (int c = 0; c < 2; c++) {
            double d = (ca[c] - pa[l][c]) - (cb[c] - pb[l][c]);
            bal += d * d;
        }
    }
    ee /= m->nee;
    bal /= m->nee * m->height;
    double com = 0, dcv = 0;
    for (int c = 0; c < 3; c++) { com += (ca[c] - cb[c]) * (ca[c] - cb[c]); dcv += (cva[c] - cvb[c]) * (cva[c] - cvb[c]); }
    com += 0.1 * dcv;
    out[1] = pose; out[2] = root; out[3] = ee; out[4] = bal; out[5] = com;
    out[0] = m->cw[0] * pose + m->cw[1] * root + m->cw[2] * ee + m->cw[3] * bal + m->cw[4] * com;
}

/* ------------------------------------------------------------------ one SamCon window for a slice of samples
 * /root/reference/samcon/core/samcon.py:28-41: sample k restores elite k / (S/E), steps, is scored. */
void oc_window_cache(const oc_model *m, const double *parents /* [E,132] */, const double *actions /* [S,68] */,
                     const double *target, int k0, int k1, int children, int nsteps, double *out_state, double *out_cost,
                     double *out_info, const double *parent_cache /* [E, 2 maxc] or NULL */, double *out_cache /* [S, 2 maxc] or NULL */) {
    int sd = 13 + 7 * m->nj, ad = 4 * m->nj, cw = 2 * m->maxc;
    for (int k = k0; k < k1; k++) {
        double s[256], c[6], cache[2 * MAXCC];
        memcpy(s, parents + (size_t)(k / children) * sd, sizeof(double) * sd);
        if (parent_cache) memcpy(cache, parent_cache + (size_t)(k / children) * cw, sizeof(double) * cw);
        else for (int i = 0; i < cw; i++) cache[i] = i < m->maxc ? -1.0 : 0.0;
        oc_step_cache(m, s, actions + (size_t)k * ad, nsteps, cache);
        if (out_cache) memcpy(out_cache + (size_t)k * cw, cache, sizeof(double) * cw);
        oc_cost(m, s, target, c);
        memcpy(out_state + (size_t)k * sd, s, sizeof(double) * sd);
        out_cost[k] = c[0];
        memcpy(out_info + (size_t)k * 5, c + 1, 40);
    }
}

void oc_window(const oc_model *m, const double *parents, const double *actions, const double *target, int k0, int k1,
               int children, int nsteps, double *out_state, double *out_cost, double *out_info) {
    oc_window_cache(m, parents, actions, target, k0, k1, children, nsteps, out_state, out_cost, out_info, NULL, NULL);
}

/* elite selection: /root/reference/samcon/core/trajectory.py:108-112, ties broken by index */
void oc_topk(const double *cost, int n, int k, int *idx) {
    char *used = (char *)calloc(n, 1);
    for (int r = 0; r < k; r++) {
        int best = -1;
        for (int i = 0; i < n; i++) if (!used[i] && (best < 0 || cost[i] < cost[best])) best = i;
        used[best] = 1; idx[r] = best;
    }
    free(used);
}

/* ------------------------------------------------------------------ diagnostics for tests */
void oc_mass_matrix(const oc_model *m, const double *state, int set, double *M) {
    oc_state st; oc_kin k;
    unpack(m, state, &st); fk(m, &st, &k); crba(m, set, &k, M);
}
void oc_bias(const oc_model *m, const double *state, int set, int grav, const double *qdd, double *tau) {
    oc_state st; oc_kin k;
    unpack(m, state, &st); fk(m, &st, &k); rnea(m, set, &st, &k, qdd, grav, tau);
}
void oc_gen_vel(const oc_model *m, const double *state, double *v) {
    oc_state st; oc_kin k;
    unpack(m, state, &st); fk(m, &st, &k); gen_vel(m, &st, &k, v);
}
void oc_pd_torque(const oc_model *m, const double *state, const double *action, double *tau) {
    oc_state st; oc_kin k;
    unpack(m, state, &st); fk(m, &st, &k); stable_pd(m, &st, &k, action, tau);
}
void oc_link_coms(const oc_model *m, const double *state, double *pos, double *vel, double *com /* 6 */) {
    oc_state st;
    unpack(m, state, &st);
    link_coms(m, &st, (double(*)[3])pos, (double(*)[3])vel, com, com + 3);
}
int oc_contacts(const oc_model *m, const double *state, double *out /* [maxc,5]: link,x,y,z,dist */) {
    oc_state st; oc_kin k; oc_contact con[MAXCAND];
    unpack(m, state, &st); fk(m, &st, &k);
    int n = collide(m, &k, con);
    for (int i = 0; i < n; i++) {
        out[5 * i] = con[i].link; memcpy(out + 5 * i + 1, con[i].pos, 24); out[5 * i + 4] = con[i].dist;
    }
    return n;
}
int oc_contacts_full(const oc_model *m, const double *state, double *out /* [maxc,12]: link,link2,pos,pos2,n,dist */) {
    oc_state st; oc_kin k; oc_contact con[MAXCAND];
    unpack(m, state, &st); fk(m, &st, &k);
    int n = collide(m, &k, con);
    for (int i = 0; i < n; i++) {
        double *o = out + 12 * i;
        o[0] = con[i].link; o[1] = con[i].link2; memcpy(o + 2, con[i].pos, 24); memcpy(o + 5, con[i].pos2, 24);
        memcpy(o + 8, con[i].n, 24); o[11] = con[i].dist;
    }
    return n;
}
int oc_num_pairs(const oc_model *m) { return m->npair; }
double oc_box_segment(const double *h, const double *p0, const double *p1, double *q, double *p, int *axis) {
    double sgn;
    return box_segment(h, p0, p1, q, p, axis, &sgn);
}
/* free dynamics with given generalised forces and no stable PD: used by conservation tests */
void oc_step_torque(const oc_model *m, double *state, const double *tau, int nsub_total) {
    oc_state st; oc_kin k;
    unpack(m, state, &st);
    for (int u = 0; u < nsub_total; u++) { fk(m, &st, &k); substep(m, &st, &k, tau, m->dt / m->nsub); }
    pack(m, &st, state);
}

// sc_core.cuh -- per-sample humanoid rollout, written as "lane stages".
//
// Mapping: one sample is owned by a group of LPS = 8 lanes; a warp carries a TILE of 4 samples.
// lane = slot*4 + s  (s = sample in tile 0..3, slot = 0..7).  Bodies are numbered level by level and no tree
// level is wider than 4, so in every tree sweep the group's lane `slot` owns body level_start[d]+slot.
// All per-sample working data lives in shared memory, interleaved so that word w of sample s sits at
// sm[w*4+s] (the 4 samples of a tile hit 4 different banks; the <=4 bodies of a level have consecutive ids).
// A "stage" is a piece of straight-line per-lane code followed by a warp barrier; no per-lane value lives
// across a barrier.  That discipline lets the same source be compiled for the host as a lane loop
// (tests/emu, debugging aid only -- the package never loads it) and for sm_100a with __syncwarp().
//
// Arithmetic follows the reference call chain env.step -> actuate -> stepSimulation -> cost
// (/root/reference/envs/humanoid_tracking_env.py:184-200, humanoid_stable_pd.py:124-135) as restated in
// DESIGN.md §2; algorithms: articulated-body recursion specialised to spherical joints about the joint
// centre (S=[1;0]), stable PD as an ABA with joint armature dt*Kd, contact-space Gauss-Seidel.
// All spatial quantities are expressed in WORLD-ALIGNED axes about the owning body's origin (Pluecker transforms
// between bodies are pure shifts by r = p_child - p_parent, no rotations): a spherical joint's motion subspace is
// the whole angular 3-space in any basis and the armature dt*Kd is isotropic, so only the constant rigid-body
// inertia has to be rotated into world axes, never the 6x6 articulated inertias.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "../../include/samcon_b200.h"

#if defined(__CUDACC__)
#define SC_DEV __device__ __forceinline__
// phase-level functions; measured on B200: making them real calls (__noinline__, 14.4k -> 13.1k SASS instructions)
// was 3-5 % slower than inlining them into the step loop, also when only the ones with several call sites (fk_sweep,
// aba) are shared
#define SC_PHASE __device__ __forceinline__
#define SC_CALL __device__ __noinline__
#define SC_SYNC() __syncwarp()
// CTA-wide phase alignment.  Not needed for correctness (warps never share data): it keeps the warps of a CTA in
// the same code region, because the step loop is ~170 KB of SASS and 8 warps at 8 different places in it starve
// on instruction fetch (ncu: stalled_no_instruction 7.7 per issue on the walk workload without it, 0.2 with the
// warps aligned).  Every warp of the CTA must execute the same number of these, see rollout_tile().
// (named barriers over groups of 5 / 2 warps instead of the whole CTA: 39.8 / 67 ms against 32.7 ms)
#define SC_CTA_SYNC() __syncthreads()
#ifndef SC_BAR_MASK
#define SC_BAR_MASK 31
#endif
#define SC_BAR(bit) do { if (SC_BAR_MASK & (bit)) SC_CTA_SYNC(); } while (0)
// the step loop's barriers are also switched per kernel instance (template parameter BM of rollout_tile; a run-time mask
// loses what dropping a barrier gains -- the compiler schedules across a barrier that is not there): which ones pay
// depends on how many warps the CTA carries.  Bit 16, behind the shared assembly, is the only one needed for correctness.
#define SC_BARL(bit) do { if ((SC_BAR_MASK & (bit)) && ((BM | 16) & (bit))) SC_CTA_SYNC(); } while (0)
// Contact-matrix assembly shared between the warps of a CTA (assembly_shared below): a warp that has finished its own
// tile's columns takes column calls of the tiles that have not, instead of standing at the barrier after the assembly.
#ifndef SC_SHARE_ASSEMBLY
#define SC_SHARE_ASSEMBLY 1
#endif
#ifndef SC_POLL_NS
#define SC_POLL_NS 100
#endif
#else
#define SC_DEV static inline
#define SC_PHASE static
#define SC_CALL static
#define SC_SYNC() ((void)0)
#define SC_CTA_SYNC() ((void)0)
#define SC_BAR(bit) ((void)0)
#define SC_BARL(bit) ((void)0)
#endif

// Optional phase timing (developer builds only: -DSC_PROFILE, tools/phase_profile.py).  SC_T(i) charges the cycles
// since the previous mark to phase i of this warp.
#if defined(SC_PROFILE) && defined(__CUDACC__)
#define SC_NPHASE 40
#define SC_NTRACE 2048            // behind the counters: [0] events written, [8 + 4 k ..] event k (assembly_shared)
namespace sc { __device__ long long *sc_trace = nullptr; }
#define SC_T(i) do { long long t_ = clock64(); sc_prof_acc[i] += t_ - sc_prof_last; sc_prof_last = t_; } while (0)
#define SC_PROF_ARGS , long long *sc_prof_acc, long long &sc_prof_last
#define SC_PROF_PASS , sc_prof_acc, sc_prof_last
#else
#define SC_T(i) ((void)0)
#define SC_PROF_ARGS
#define SC_PROF_PASS
#endif

namespace sc {

constexpr int NB = SC_MAX_BODIES;
constexpr int MAXLEV = SC_MAX_LEVELS;
constexpr int MAXC = SC_MAX_CONTACTS;
constexpr int NR = 3 * MAXC;
#ifndef SC_TILE
#define SC_TILE 4
#endif
constexpr int TILE = SC_TILE;        // samples per warp
constexpr int LPS = 32 / SC_TILE;    // lanes per sample
static_assert(TILE * LPS == 32, "one tile per warp");

// self-collision tables (part of Tables, in shared memory)
constexpr int NPRIM = SC_MAX_PRIMS, NPAIR = SC_MAX_PAIRS, NBOX = SC_MAX_BOXES;
constexpr int PAIRS_PER_LANE_MAX = (NPAIR + LPS - 1) / LPS, PRIMS_PER_LANE_MAX = (NPRIM + LPS - 1) / LPS;
struct SelfTables {
    int nprim, npair, prims_per_lane, pairs_per_lane;
    float prim[NPRIM * 8];              // sphere-swept segment in the body frame: p0, p1, radius, bounding radius about the midpoint
    float pair_r2[NPAIR];               // (bounding radius a + bounding radius b + contact threshold)^2
    unsigned char pair_a[NPAIR], pair_b[NPAIR], prim_body[NPRIM];
    int nbox;                           // exact boxes (0: every box is its capsule stand-in `prim`)
    signed char prim_box[NPRIM];        // box index of the primitive or -1
    float box[NBOX * 15];               // body frame: centre (3), axes R row major with v_body = R v_box (9), half extents (3)
    float box_core[NBOX * 7];           // the box's long axis as a segment (p0, p1, body frame) and the radius of the capsule
                                        // around it that CONTAINS the box: a cheap lower bound of any distance to the box
};

// ---- derived model tables (copied to shared memory once per CTA) ------------------------------------------
struct Tables {
    int nb, nlev, ncand, ncp, nee, niter, nsub, maxc, nj;
    int parent[NB], depth[NB], jidx[NB], jbody[NB], nchild[NB], child[NB][4], level_start[MAXLEV + 1], anc_mask[NB];
    int icslot[NB];        // O_IC slot of body b: (depth & 1) * ICW + position in its level
    int anc[NB][MAXLEV];   // anc[b][d] = ancestor of b at tree depth d (d <= depth[b]; anc[b][depth[b]] = b), else -1
    int cand_per_lane;
    int cand_body[SC_MAX_CAND], cp_body[SC_MAX_CPOINTS], ee[4];
    float jorg[NB * 3], idyn[NB * 10], ipd[NB * 10], kp[NB], kd[NB], maxf[NB], cand[SC_MAX_CAND * 4];
    float cp_off[SC_MAX_CPOINTS * 3], cp_mass[SC_MAX_CPOINTS], jw[NB], cw[5];
    float height, dt, h, g[3], mu, mu_self, erp, cthresh, maxvel, inv_total_mass, warmstart;
    SelfTables self;
};
// The same tables under two static types: the rollout is compiled twice, with and without the exact box narrow phase, so
// that the default (boxes as their capsule stand-ins) does not carry -- or fetch -- the code of a routine it never runs.
struct TablesCapsule : Tables { static constexpr bool BOX = false; };
struct TablesBox : Tables { static constexpr bool BOX = true; };


// ---- per-sample shared-memory layout (32-bit words) --------------------------------------------------------
// Regions marked (b>=1) have no slot for the root: their offset is pulled back by one stride so that
// `O_X + stride * b` still addresses body b; the root's would-be slot overlaps the tail of the region before it and is
// never touched (every per-body loop below treats b == 0 separately).
constexpr int ICW = 4;                            // articulated-inertia slots per tree level (max level width)
constexpr int O_P0 = 0;                           // root position (3)
constexpr int O_Q = O_P0 + 3;                     // quaternions xyzw: body 0 = root orientation, b>0 = joint rotation
constexpr int O_VB = O_Q + 4 * NB;                // base world velocity: angular(3), linear(3)
constexpr int O_W = O_VB + 6 - 3;                 // joint angular velocity in child frame (3, b>=1)
constexpr int O_RW = O_W + 3 * NB;                // world rotation per body (9)
constexpr int O_PW = O_RW + 9 * NB;               // world position of body origin (3)
constexpr int O_RJ = O_PW + 3 * NB - 3;           // r = p_b - p_parent in world axes (3, b>=1)
constexpr int O_TW = O_RJ + 3 * NB;               // body twist, world axes: omega(3), velocity of the body origin(3)
constexpr int O_WJ = O_TW + 6 * NB - 3;           // joint angular velocity in world axes, R_b * w (3, b>=1)
constexpr int O_TAU = O_WJ + 3 * NB - 3;          // child-frame joint torque held over the sub-steps; during the
                                                  // stable-PD solve: Kp*err - Kd*w (3, b>=1)
constexpr int O_UA = O_TAU + 3 * NB - 6;          // articulated inertia (world axes), angular block (sym 6, b>=1): written and read
                                                  // by the stable-PD pass only (the dynamics pass has K A = 1); its words carry the
                                                  // contact list during the sub-steps, see O_CB below
constexpr int O_UB = O_UA + 6 * NB - 9;           // articulated inertia, linear-from-angular block (9, b>=1)
constexpr int O_DI = O_UB + 9 * NB - 6;           // (S^T IA S + armature)^-1 (sym 6, b>=1)
constexpr int O_I0L = O_DI + 6 * NB;              // Cholesky factor of the base articulated inertia (21)
// scratch block: O_IC .. O_SCR_END is contiguous; the contact-space matrix O_AM overlays all of it while none of
// its members is live (between the end of the dynamics ABA and the start of the impulse application)
constexpr int O_IC = O_I0L + 21;                  // articulated-inertia contribution to the parent, A(6) B(9) C(6):
                                                  // one slot per body of two adjacent tree levels (T.icslot)
constexpr int O_PC = O_IC + 21 * 2 * ICW;         // bias-force contribution to the parent (6); later: accelerations
constexpr int O_UU = O_PC + 6 * NB - 3;           // u = tau - S^T pA (3, b>=1); contact apply: -S^T pa
constexpr int O_SCR_END = O_UU + 3 * NB;
// contact cache: survives from one sub-step to the next, across the stable-PD pass and (through parent_cache /
// out_cache) from one window to the next -- what Bullet keeps in its persistent manifolds and saveBullet stores
constexpr int O_PCU = O_SCR_END;                  // candidate id of contact j (int): the current list once collide's cache stage ran
constexpr int O_PLAM = O_PCU + MAXC;              // its normal impulse: warm-start value before the sweep, result after it
constexpr int O_PNC = O_PLAM + MAXC;              // number of cached contacts (int)
constexpr int O_OVF = O_PNC + 1;                  // sub-steps of this sample with more candidate hits than max_contacts (int)
constexpr int O_END_DEV = O_OVF + 1;
// contact list of the current sub-step, in the words of O_UA (dead between two stable-PD passes)
constexpr int O_CB = O_UA + 6;                    // contact bodies (int): body a | (body b + 1) << 8, b = -1: the ground
constexpr int O_CU = O_CB + MAXC;                 // candidate id (int): primitive pair u < npair, else npair + ground point
constexpr int O_CR = O_CU + MAXC;                 // contact point relative to the body origin, world axes (3)
constexpr int O_CD = O_CR + 3 * MAXC;             // signed distance
constexpr int O_NC = O_CD + MAXC;                 // number of contacts (int)
constexpr int O_TM = O_NC + 1;                    // mask of bodies on a root path of some contact (int), | TM_TWO_BODY
constexpr int TM_TWO_BODY = 1 << 30;              // some contact of the sample is link-vs-link
constexpr int O_RHS = O_TM + 1;                   // right-hand sides; the sweep overwrites them with the impulses:
constexpr int O_LAM = O_RHS;                      // a lane reads the rows it owns before the sweep and writes the same rows after it
constexpr int O_DM = O_RHS + 3 * MAXC;            // mask of the bodies that carry a contact themselves (int)
static_assert(O_DM < O_UB, "the contact list fits into the angular-block words the dynamics pass leaves free");
static_assert(O_RHS + NR <= O_UB + 9, "the contact list must fit in the words of O_UA");
constexpr int O_AM = O_IC;                        // contact-space matrix, packed lower triangle (contact solve)
constexpr int O_ZC = O_IC;                        // collide: signed distance per candidate (pairs, then ground points) ...
constexpr int O_CNT = O_ZC + SC_MAX_CAND + NPAIR; // ... hits per lane (ground, pairs) ...
constexpr int O_PCEN = O_CNT + 2 * LPS;          // ... and world centres of the self-collision primitives (3)
constexpr int HL_MAX = 32;                        // hits (distance, candidate) compacted over O_PCEN when more than maxc
constexpr int O_HL = O_PCEN;
static_assert(2 * HL_MAX <= 3 * NPRIM, "hit list must fit in PCEN");
// second body of a contact: live from collide to the end of the contact solve, while O_WJ is dead
constexpr int O_CR2 = O_WJ + 3;                   // contact point relative to body b's origin, world axes (3)
constexpr int O_CN = O_CR2 + 3 * MAXC;            // contact normal, from body b (or the ground) to body a (3)
constexpr int O_FEAT = O_RHS;                     // end of window: com(3) comvel(3) ee(4x3)
static_assert(NR * (NR + 1) / 2 <= O_SCR_END - O_IC, "AM must fit in the scratch block");
static_assert(O_PCEN + 3 * NPRIM <= O_SCR_END, "collide scratch must fit in the scratch block");
static_assert(O_CN + 3 * MAXC <= O_WJ + 3 * NB, "second-body contact data must fit in WJ");
static_assert(18 <= NR, "FEAT must fit in RHS");
#if defined(__CUDACC__)
constexpr int O_END = O_END_DEV;
#else
// host emulation of the Gauss-Seidel sweep keeps its per-row state in (emulated) shared memory
constexpr int O_RES = O_END_DEV;
constexpr int O_DGI = O_RES + NR;
constexpr int O_DEL = O_DGI + NR;         // pending impulse change, double buffered (2)
constexpr int O_END = O_DEL + 2;
#endif
constexpr int WORDS_PER_SAMPLE = O_END;
constexpr int WORDS_PER_TILE = WORDS_PER_SAMPLE * TILE;

#define F(i) S[(i) * sc::TILE]

SC_DEV int f2i(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
SC_DEV float i2f(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}
SC_DEV float rsq(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}

// ---- small vector algebra in registers -------------------------------------------------------------------
struct V3 { float x, y, z; };
struct M3 { float a[9]; };            // row major
struct S3 { float xx, xy, xz, yy, yz, zz; };

SC_DEV V3 mk(float x, float y, float z) { V3 v = {x, y, z}; return v; }
SC_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
SC_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
SC_DEV V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
SC_DEV V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
SC_DEV float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SC_DEV V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
SC_DEV V3 mul(const M3 &R, V3 v) {
    return mk(R.a[0] * v.x + R.a[1] * v.y + R.a[2] * v.z, R.a[3] * v.x + R.a[4] * v.y + R.a[5] * v.z,
              R.a[6] * v.x + R.a[7] * v.y + R.a[8] * v.z);
}
SC_DEV V3 mulT(const M3 &R, V3 v) {
    return mk(R.a[0] * v.x + R.a[3] * v.y + R.a[6] * v.z, R.a[1] * v.x + R.a[4] * v.y + R.a[7] * v.z,
              R.a[2] * v.x + R.a[5] * v.y + R.a[8] * v.z);
}
SC_DEV V3 mul(const S3 &A, V3 v) {
    return mk(A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z,
              A.xz * v.x + A.yz * v.y + A.zz * v.z);
}
SC_DEV M3 mul(const M3 &A, const M3 &B) {
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
    return C;
}
SC_DEV M3 mulBt(const M3 &A, const M3 &B) {  // A * B^T
    M3 C;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C.a[3 * i + j] = A.a[3 * i] * B.a[3 * j] + A.a[3 * i + 1] * B.a[3 * j + 1] + A.a[3 * i + 2] * B.a[3 * j + 2];
    return C;
}
SC_DEV M3 tom(const S3 &A) { M3 M = {{A.xx, A.xy, A.xz, A.xy, A.yy, A.yz, A.xz, A.yz, A.zz}}; return M; }
SC_DEV S3 upper(const M3 &M) { S3 A = {M.a[0], M.a[1], M.a[2], M.a[4], M.a[5], M.a[8]}; return A; }
SC_DEV V3 row(const M3 &M, int i) { return mk(M.a[3 * i], M.a[3 * i + 1], M.a[3 * i + 2]); }
SC_DEV V3 col(const M3 &M, int j) { return mk(M.a[j], M.a[3 + j], M.a[6 + j]); }
SC_DEV void setrow(M3 &M, int i, V3 v) { M.a[3 * i] = v.x; M.a[3 * i + 1] = v.y; M.a[3 * i + 2] = v.z; }
SC_DEV void setcol(M3 &M, int j, V3 v) { M.a[j] = v.x; M.a[3 + j] = v.y; M.a[6 + j] = v.z; }
SC_DEV S3 inv(const S3 &A) {
    float c0 = A.yy * A.zz - A.yz * A.yz, c1 = A.xz * A.yz - A.xy * A.zz, c2 = A.xy * A.yz - A.xz * A.yy;
    float id = 1.0f / (A.xx * c0 + A.xy * c1 + A.xz * c2);
    S3 K = {c0 * id, c1 * id, c2 * id, (A.xx * A.zz - A.xz * A.xz) * id, (A.xz * A.xy - A.xx * A.yz) * id,
            (A.xx * A.yy - A.xy * A.xy) * id};
    return K;
}
SC_DEV M3 quat2R(float x, float y, float z, float w) {
    M3 R = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
             2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    return R;
}
// R A R^T for a symmetric A
SC_DEV S3 rotS(const M3 &R, const S3 &A) {
    const V3 m0 = mul(A, row(R, 0)), m1 = mul(A, row(R, 1)), m2 = mul(A, row(R, 2));   // columns of A R^T
    S3 O = {dot(row(R, 0), m0), dot(row(R, 0), m1), dot(row(R, 0), m2), dot(row(R, 1), m1), dot(row(R, 1), m2),
            dot(row(R, 2), m2)};
    return O;
}

// ---- shared-memory load/store helpers ----------------------------------------------------------------------
SC_DEV V3 ld3(const float *S, int o) { return mk(F(o), F(o + 1), F(o + 2)); }
SC_DEV void st3(float *S, int o, V3 v) { F(o) = v.x; F(o + 1) = v.y; F(o + 2) = v.z; }
SC_DEV M3 ldM(const float *S, int o) {
    M3 M;
#pragma unroll
    for (int i = 0; i < 9; i++) M.a[i] = F(o + i);
    return M;
}
SC_DEV void stM(float *S, int o, const M3 &M) {
#pragma unroll
    for (int i = 0; i < 9; i++) F(o + i) = M.a[i];
}
SC_DEV S3 ldS(const float *S, int o) { S3 A = {F(o), F(o + 1), F(o + 2), F(o + 3), F(o + 4), F(o + 5)}; return A; }
SC_DEV void stS(float *S, int o, const S3 &A) {
    F(o) = A.xx; F(o + 1) = A.xy; F(o + 2) = A.xz; F(o + 3) = A.yy; F(o + 4) = A.yz; F(o + 5) = A.zz;
}
SC_DEV M3 ldRj(const float *S, int b) { return quat2R(F(O_Q + 4 * b), F(O_Q + 4 * b + 1), F(O_Q + 4 * b + 2), F(O_Q + 4 * b + 3)); }
SC_DEV V3 tv3(const float *t, int i) { return mk(t[3 * i], t[3 * i + 1], t[3 * i + 2]); }

// 6x6 SPD solve, packed lower triangle, diagonal stored inverted
SC_DEV constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }
SC_DEV void chol6(float *a) {
#pragma unroll
    for (int j = 0; j < 6; j++) {
        float d = a[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= a[tri(j, k)] * a[tri(j, k)];
        float id = rsq(d);
        a[tri(j, j)] = id;
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float s = a[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) s -= a[tri(i, k)] * a[tri(j, k)];
            a[tri(i, j)] = s * id;
        }
    }
}
SC_DEV void solve6(const float *L, float *x) {
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float s = x[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[tri(i, k)] * x[k];
        x[i] = s * L[tri(i, i)];
    }
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        float s = x[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) s -= L[tri(k, i)] * x[k];
        x[i] = s * L[tri(i, i)];
    }
}

// max over the samples of a tile of an int field; call between stages only
SC_DEV int tile_max_int(const float *sm, int off) {
#if defined(__CUDACC__)
    int v = f2i(sm[off * TILE + (threadIdx.x & (TILE - 1))]);
#pragma unroll
    for (int o = 1; o < TILE; o <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
#else
    int v = f2i(sm[off * TILE]);
    for (int s = 1; s < TILE; s++) { int u = f2i(sm[off * TILE + s]); v = u > v ? u : v; }
    return v;
#endif
}

#define SC_STAGE for (int lane = L0; lane < L1; ++lane)
#define SC_LANE float *S = sm + (lane & (sc::TILE - 1)); const int slot = lane / sc::TILE; (void)slot; (void)S

// ============================================================================================================
// kinematics
// ============================================================================================================
SC_DEV void fk_body(float *S, const Tables &T, int b, bool full) {
    if (b == 0) {
        if (full) { stM(S, O_RW, ldRj(S, 0)); st3(S, O_PW, ld3(S, O_P0)); }
        st3(S, O_TW, ld3(S, O_VB));
        st3(S, O_TW + 3, ld3(S, O_VB + 3));
    } else {
        const int p = T.parent[b];
        M3 R;
        V3 r;
        if (full) {
            const M3 Rp = ldM(S, O_RW + 9 * p);
            R = mul(Rp, ldRj(S, b));
            r = mul(Rp, tv3(T.jorg, b));
            stM(S, O_RW + 9 * b, R);
            st3(S, O_RJ + 3 * b, r);
            st3(S, O_PW + 3 * b, ld3(S, O_PW + 3 * p) + r);
        } else {
            R = ldM(S, O_RW + 9 * b);
            r = ld3(S, O_RJ + 3 * b);
        }
        const V3 wjw = mul(R, ld3(S, O_W + 3 * b));
        const V3 wp = ld3(S, O_TW + 6 * p), vp = ld3(S, O_TW + 6 * p + 3);
        if (full) st3(S, O_WJ + 3 * b, wjw);      // the velocity-only pass runs while O_WJ holds contact data
        st3(S, O_TW + 6 * b, wp + wjw);
        st3(S, O_TW + 6 * b + 3, vp + cross(wp, r));
    }
}

template <class TT>
SC_PHASE void fk_sweep(float *sm, const TT &T, int L0, int L1, bool full) {
    for (int d = 0; d < T.nlev; d++) {
        SC_STAGE {
            SC_LANE;
            int b = T.level_start[d] + slot;
            if (b < T.level_start[d + 1]) fk_body(S, T, b, full);
        }
        SC_SYNC();
    }
}

// ============================================================================================================
// articulated-body algorithm (PD = stable-PD variant: pd inertia set, armature dt*Kd, O_TAU holds Kp err - Kd w)
// ============================================================================================================
SC_DEV void aba_backward_body(float *S, const Tables &T, int b, const bool PD) {
    const float *I = (PD ? T.ipd : T.idyn) + 10 * b;
    const float m = I[0];
    const M3 R = ldM(S, O_RW + 9 * b);
    // rigid-body inertia about the body origin, rotated into world axes
    const V3 hh = mul(R, mk(I[1], I[2], I[3]));
    const S3 Ab = {I[4], I[5], I[6], I[7], I[8], I[9]};
    S3 A = rotS(R, Ab);
    M3 B = {{0, hh.z, -hh.y, -hh.z, 0, hh.x, hh.y, -hh.x, 0}};   // -skew(h)
    S3 C = {m, 0, 0, m, 0, m};
    const V3 w = ld3(S, O_TW + 6 * b), v = ld3(S, O_TW + 6 * b + 3);
    // own bias force v x* (I v)
    V3 n0 = mul(A, w) + cross(hh, v);
    V3 f0 = m * v - cross(hh, w);
    V3 pn = cross(w, n0) + cross(v, f0);
    V3 pf = cross(w, f0);
    for (int k = 0; k < T.nchild[b]; k++) {
        const int c = T.child[b][k];
        const int o = O_IC + 21 * T.icslot[c];
        A.xx += F(o); A.xy += F(o + 1); A.xz += F(o + 2); A.yy += F(o + 3); A.yz += F(o + 4); A.zz += F(o + 5);
#pragma unroll
        for (int i = 0; i < 9; i++) B.a[i] += F(o + 6 + i);
        C.xx += F(o + 15); C.xy += F(o + 16); C.xz += F(o + 17); C.yy += F(o + 18); C.yz += F(o + 19); C.zz += F(o + 20);
        pn = pn + ld3(S, O_PC + 6 * c);
        pf = pf + ld3(S, O_PC + 6 * c + 3);
    }
    if (b == 0) {
        // floating base: a0 = -IA^-1 pA in the free-fall frame; gravity is added when velocities are updated
        float L[21] = {A.xx, A.xy, A.yy, A.xz, A.yz, A.zz,
                       B.a[0], B.a[1], B.a[2], C.xx,
                       B.a[3], B.a[4], B.a[5], C.xy, C.yy,
                       B.a[6], B.a[7], B.a[8], C.xz, C.yz, C.zz};
        chol6(L);
        float x[6] = {-pn.x, -pn.y, -pn.z, -pf.x, -pf.y, -pf.z};
        solve6(L, x);
#pragma unroll
        for (int i = 0; i < 21; i++) F(O_I0L + i) = L[i];
#pragma unroll
        for (int i = 0; i < 6; i++) F(O_PC + i) = x[i];
        if (!PD) {
            // base velocity update: w += h alpha ; v += h (a + w x v + g)   (spatial -> classical acceleration).
            // The change of the body twist is handed to the children (O_IC slot, free in the outward sweep), so that
            // the sweep leaves O_TW at the new velocities and no separate velocity pass is needed before the contacts.
            const V3 a = mk(x[3], x[4], x[5]) + cross(w, v);
            const V3 dw = T.h * mk(x[0], x[1], x[2]), dv = T.h * mk(a.x + T.g[0], a.y + T.g[1], a.z + T.g[2]);
            st3(S, O_VB, ld3(S, O_VB) + dw); st3(S, O_VB + 3, ld3(S, O_VB + 3) + dv);
            st3(S, O_TW, w + dw); st3(S, O_TW + 3, v + dv);
            st3(S, O_IC + 21 * T.icslot[0], dw); st3(S, O_IC + 21 * T.icslot[0] + 3, dv);
        }
        return;
    }
    const float arm = PD ? T.dt * T.kd[b] : 0.0f;
    S3 D = A; D.xx += arm; D.yy += arm; D.zz += arm;
    const S3 K = inv(D);
    const V3 u = mul(R, ld3(S, O_TAU + 3 * b)) - pn;      // child-frame torque (PD: Kp err - Kd w) in world axes
    const V3 wj = ld3(S, O_WJ + 3 * b);
    const V3 cw = cross(w, wj), cv = cross(v, wj);
    if (PD) stS(S, O_UA + 6 * b, A);      // the dynamics pass never reads it back (K A = 1): its words hold the contact list
    stM(S, O_UB + 9 * b, B); stS(S, O_DI + 6 * b, K); st3(S, O_UU + 3 * b, u);
    // Articulated inertia seen by the parent, IA - IA S (S^T IA S + arm)^-1 S^T IA with S = [1; 0] and K = (A + arm 1)^-1.
    // K commutes with A and A K = 1 - arm K, so with P = arm K:
    //   Aa = A - A K A = arm (1 - P),   Ba = B - B K A = arm B K,   Ca = C - B K B^T,   A K u = u - P u
    // (arm = 0, the dynamics pass: a spherical joint passes no moment but its torque -- Aa = Ba = 0).
    const M3 BK = mul(B, tom(K));
    M3 Ca;
    {
        // Ca = C - BK B^T is symmetric: six dot products
        const float c00 = dot(row(BK, 0), row(B, 0)), c01 = dot(row(BK, 0), row(B, 1)), c02 = dot(row(BK, 0), row(B, 2));
        const float c11 = dot(row(BK, 1), row(B, 1)), c12 = dot(row(BK, 1), row(B, 2)), c22 = dot(row(BK, 2), row(B, 2));
        const S3 Cs = {C.xx - c00, C.xy - c01, C.xz - c02, C.yy - c11, C.yz - c12, C.zz - c22};
        Ca = tom(Cs);
    }
    // shift of the reference point to the parent's origin (r = p_b - p_parent); axes are shared, nothing to rotate:
    // Bp = Ba - Ca rx,  Ap = Aa - Ba^T rx + rx Ba - rx Ca rx = Aa - Ba^T rx + rx Bp
    const V3 r = ld3(S, O_RJ + 3 * b);
    M3 CR, Bp, Ap, RB;
#pragma unroll
    for (int i = 0; i < 3; i++) setrow(CR, i, cross(row(Ca, i), r));              // Ca * rx
    V3 pan, paf = pf + mul(Ca, cv) + mul(BK, u);
    if (PD) {
        const S3 P = {arm * K.xx, arm * K.xy, arm * K.xz, arm * K.yy, arm * K.yz, arm * K.zz};
        const S3 Aas = {arm - arm * P.xx, -arm * P.xy, -arm * P.xz, arm - arm * P.yy, -arm * P.yz, arm - arm * P.zz};
        const M3 Aa = tom(Aas);
        M3 Ba, BtR;
#pragma unroll
        for (int i = 0; i < 9; i++) Ba.a[i] = arm * BK.a[i];
        pan = pn + mul(Aas, cw) + mulT(Ba, cv) + (u - mul(P, u));
        paf = paf + mul(Ba, cw);
#pragma unroll
        for (int i = 0; i < 9; i++) Bp.a[i] = Ba.a[i] - CR.a[i];
#pragma unroll
        for (int i = 0; i < 3; i++) setrow(BtR, i, cross(col(Ba, i), r));         // Ba^T * rx
#pragma unroll
        for (int j = 0; j < 3; j++) setcol(RB, j, cross(r, col(Bp, j)));          // rx * Bp
#pragma unroll
        for (int i = 0; i < 9; i++) Ap.a[i] = Aa.a[i] - BtR.a[i] + RB.a[i];
    } else {
        // no armature: Aa = Ba = 0, the joint passes on its torque and nothing else of the moment
        pan = pn + u;
#pragma unroll
        for (int i = 0; i < 9; i++) Bp.a[i] = -CR.a[i];
#pragma unroll
        for (int j = 0; j < 3; j++) setcol(Ap, j, cross(r, col(Bp, j)));          // rx * Bp
    }
    const int o = O_IC + 21 * T.icslot[b];
    stS(S, o, upper(Ap)); stM(S, o + 6, Bp); stS(S, o + 15, upper(Ca));
    st3(S, O_PC + 6 * b, pan + cross(r, paf));
    st3(S, O_PC + 6 * b + 3, paf);
}

SC_DEV void aba_forward_body(float *S, const Tables &T, int b, const bool PD) {
    const int p = T.parent[b];
    const V3 r = ld3(S, O_RJ + 3 * b);
    const V3 alp = ld3(S, O_PC + 6 * p), ap = ld3(S, O_PC + 6 * p + 3);
    const V3 w = ld3(S, O_TW + 6 * b), v = ld3(S, O_TW + 6 * b + 3), wj = ld3(S, O_WJ + 3 * b);
    const V3 aw = alp + cross(w, wj);
    const V3 al = ap + cross(alp, r) + cross(v, wj);
    const S3 K = ldS(S, O_DI + 6 * b);
    const M3 B = ldM(S, O_UB + 9 * b);
    V3 qdd;
    if (PD) qdd = mul(K, ld3(S, O_UU + 3 * b) - mul(ldS(S, O_UA + 6 * b), aw) - mulT(B, al));
    else qdd = mul(K, ld3(S, O_UU + 3 * b) - mulT(B, al)) - aw;      // no armature: K A = 1, aw + qdd = K (u - B^T al)
    st3(S, O_PC + 6 * b, aw + qdd);
    st3(S, O_PC + 6 * b + 3, al);
    const V3 qc = mulT(ldM(S, O_RW + 9 * b), qdd);        // joint acceleration in child-frame coordinates
    if (PD) {
        // tau = Kp err - Kd (w + dt qdd), clamped per axis (humanoid_stable_pd.py:52-59 max_forces)
        const float kd = T.dt * T.kd[b], mf = T.maxf[b];
        const V3 t = ld3(S, O_TAU + 3 * b) - kd * qc;
        st3(S, O_TAU + 3 * b, mk(fminf(mf, fmaxf(-mf, t.x)), fminf(mf, fmaxf(-mf, t.y)), fminf(mf, fmaxf(-mf, t.z))));
    } else {
        st3(S, O_W + 3 * b, ld3(S, O_W + 3 * b) + T.h * qc);
        // twist at the new velocities, same orientation: d omega_b = d omega_p + h qdd, d v_b = d v_p + d omega_p x r
        const int op = O_IC + 21 * T.icslot[p], ob = O_IC + 21 * T.icslot[b];
        const V3 dwp = ld3(S, op), dvp = ld3(S, op + 3);
        const V3 dw = dwp + T.h * qdd, dv = dvp + cross(dwp, r);
        st3(S, ob, dw); st3(S, ob + 3, dv);
        st3(S, O_TW + 6 * b, w + dw); st3(S, O_TW + 6 * b + 3, v + dv);
    }
}

template <class TT>
SC_PHASE void aba(float *sm, const TT &T, int L0, int L1, const bool PD) {
    for (int d = T.nlev - 1; d >= 0; d--) {
        SC_STAGE {
            SC_LANE;
            int b = T.level_start[d] + slot;
            if (b < T.level_start[d + 1]) aba_backward_body(S, T, b, PD);
        }
        SC_SYNC();
    }
    SC_BAR(256);
    for (int d = 1; d < T.nlev; d++) {
        SC_STAGE {
            SC_LANE;
            int b = T.level_start[d] + slot;
            if (b < T.level_start[d + 1]) aba_forward_body(S, T, b, PD);
        }
        SC_SYNC();
    }
}

// stable-PD proportional/derivative term: Kp * log(q_pred^-1 q_target) - Kd * w   
SC_DEV void pd_term_body(float *S, const Tables &T, int b, const float *act) {
    const float qx = F(O_Q + 4 * b), qy = F(O_Q + 4 * b + 1), qz = F(O_Q + 4 * b + 2), qw = F(O_Q + 4 * b + 3);
    const V3 w = ld3(S, O_W + 3 * b);
    const float hd = 0.5f * T.dt;
    // q_pred = normalize(q + dt/2 * q (x) (w,0))
    float px = qx + hd * (qw * w.x + qy * w.z - qz * w.y);
    float py = qy + hd * (qw * w.y - qx * w.z + qz * w.x);
    float pz = qz + hd * (qw * w.z + qx * w.y - qy * w.x);
    float pw = qw + hd * (-qx * w.x - qy * w.y - qz * w.z);
    const float *t = act + 4 * T.jidx[b];
    float tx = t[0], ty = t[1], tz = t[2], tw = t[3];
    // d = conj(p) (x) t   (scale of p and t is irrelevant for the axis, and removed for the angle by atan2)
    float dx = pw * tx - px * tw - py * tz + pz * ty;
    float dy = pw * ty + px * tz - py * tw - pz * tx;
    float dz = pw * tz - px * ty + py * tx - pz * tw;
    float dw = pw * tw + px * tx + py * ty + pz * tz;
    if (dw < 0) { dx = -dx; dy = -dy; dz = -dz; dw = -dw; }
    float s = sqrtf(dx * dx + dy * dy + dz * dz);
    float k = s > 1e-12f ? 2.0f * atan2f(s, dw) / s : 0.0f;
    const float kp = T.kp[b], kd = T.kd[b];
    st3(S, O_TAU + 3 * b, mk(kp * k * dx - kd * w.x, kp * k * dy - kd * w.y, kp * k * dz - kd * w.z));
}

// ============================================================================================================
// ground contact
// ============================================================================================================
// contact frame of a unit normal n: d[0] = n, d[1], d[2] = Bullet's btPlaneSpace1(n); the ground's n = +z gives
// t1 = -y, t2 = +x
struct Frame { V3 d[3]; };
SC_DEV Frame contact_frame(V3 n) {
    Frame fr;
    fr.d[0] = n;
    if (fabsf(n.z) > 0.70710678f) {
        const float a = n.y * n.y + n.z * n.z, k = 1.0f / sqrtf(a);
        fr.d[1] = mk(0, -n.z * k, n.y * k);
        fr.d[2] = mk(a * k, -n.x * fr.d[1].z, n.x * fr.d[1].y);
    } else {
        const float a = n.x * n.x + n.y * n.y, k = 1.0f / sqrtf(a);
        fr.d[1] = mk(-n.y * k, n.x * k, 0);
        fr.d[2] = mk(-n.z * fr.d[1].y, n.z * fr.d[1].x, a * k);
    }
    return fr;
}
// frame of contact j: ground contacts (no second body) have the constant frame +z, -y, +x
SC_DEV Frame contact_frame_of(const float *S, int j, int cb) {
    if ((cb >> 8) == 0) { Frame fr; fr.d[0] = mk(0, 0, 1); fr.d[1] = mk(0, -1, 0); fr.d[2] = mk(1, 0, 0); return fr; }
    return contact_frame(ld3(S, O_CN + 3 * j));
}
SC_DEV int cb_a(int cb) { return cb & 255; }
SC_DEV int cb_b(int cb) { return (cb >> 8) - 1; }

// Candidate points (sphere centres with a radius, box corners with radius 0) against the plane z = 0.
// Stage 1: the lanes of a sample test contiguous chunks of the ground candidates and place the self-collision
// primitives' centres.  Stage 2: broad phase over the primitive pairs (bounding spheres), narrow phase for the few
// that pass.  Stage 3: with the hit counts of the lower lanes as offset every lane appends its hits, so the contact
// list is in candidate order (primitive pairs, then ground points); only when more than `maxc` candidates are inside the
// margin does lane 0 pick the deepest serially.
constexpr int CAND_PER_LANE_MAX = (SC_MAX_CAND + LPS - 1) / LPS;
constexpr float NO_HIT = 1.0e30f;

SC_DEV float clamp01(float x) { return fminf(1.0f, fmaxf(0.0f, x)); }
// closest points of two segments (Ericson, Real-Time Collision Detection 5.1.9); parallel segments: s = 0
SC_DEV void seg_seg(V3 p1, V3 q1, V3 p2, V3 q2, V3 &c1, V3 &c2) {
    const V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    const float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    const float EPS = 1e-12f;
    float s, t;
    if (a <= EPS && e <= EPS) { s = 0; t = 0; }
    else if (a <= EPS) { s = 0; t = clamp01(f / e); }
    else {
        const float c = dot(d1, r);
        if (e <= EPS) { t = 0; s = clamp01(-c / a); }
        else {
            const float b = dot(d1, d2), den = a * e - b * b;
            s = den > 1e-6f * a * e ? clamp01((b * f - c * e) / den) : 0.0f;
            t = (b * s + f) / e;
            if (t < 0) { t = 0; s = clamp01(-c / a); }
            else if (t > 1) { t = 1; s = clamp01((b - c) / a); }
        }
    }
    c1 = p1 + s * d1; c2 = p2 + t * d2;
}
SC_DEV V3 prim_point(const float *S, const SelfTables &ST, int p, int which) {
    const int b = ST.prim_body[p];
    const float *q = ST.prim + 8 * p + 3 * which;
    return ld3(S, O_PW + 3 * b) + mul(ldM(S, O_RW + 9 * b), mk(q[0], q[1], q[2]));
}
// Closest points of a box (half extents h, its own frame) and the segment p0 + t (p1 - p0), t in [0,1] (a sphere centre
// is the segment with p1 = p0): f(t) = squared distance of the segment point to the box is convex and piecewise quadratic --
// safeguarded Newton on f' (a step solves the current piece exactly, a step that leaves the bracket becomes a bisection).
// Returns the distance of the cores, q on the box, p on the segment.  A segment that passes through the box: p = middle of
// the part inside, q = its projection on the nearest face (axis, sgn), return value = minus the depth of p below that face.
SC_DEV float box_segment(V3 h, V3 p0, V3 p1, V3 &q, V3 &p, int &axis, float &sgn) {
    // (scalars throughout: an array indexed by the exit axis would put every operand of the Newton loop into local memory)
    const V3 d = p1 - p0;
    const float dd = dot(d, d);
    float t = dd > 1e-18f ? clamp01(-dot(p0, d) / dd) : 0.0f, lo = 0.0f, hi = 1.0f;
    bool lo_seen = false, hi_seen = false;
    for (int it = 0; it < 20; it++) {
        const V3 x = p0 + t * d;
        const V3 o = mk(x.x - fmaxf(-h.x, fminf(h.x, x.x)), x.y - fmaxf(-h.y, fminf(h.y, x.y)), x.z - fmaxf(-h.z, fminf(h.z, x.z)));
        const float g1 = dot(o, d);
        const float g2 = (o.x != 0.0f ? d.x * d.x : 0.0f) + (o.y != 0.0f ? d.y * d.y : 0.0f) + (o.z != 0.0f ? d.z * d.z : 0.0f);
        if (g1 == 0.0f || g2 <= 0.0f) break;
        if (g1 > 0) { if (t <= 0.0f) break; hi = t; hi_seen = true; }
        else { if (t >= 1.0f) break; lo = t; lo_seen = true; }
        float tn = t - g1 / g2;
        if (tn <= lo) tn = lo_seen ? 0.5f * (lo + hi) : lo;
        else if (tn >= hi) tn = hi_seen ? 0.5f * (lo + hi) : hi;
        if (tn == t) break;
        t = tn;
    }
    p = p0 + t * d;
    q = mk(fmaxf(-h.x, fminf(h.x, p.x)), fmaxf(-h.y, fminf(h.y, p.y)), fmaxf(-h.z, fminf(h.z, p.z)));
    const V3 e = p - q;
    const float len2 = dot(e, e);
    axis = -1; sgn = 0;
    if (len2 > 1e-24f) return sqrtf(len2);
    // the segment enters the box: clip it against the three slabs, take the middle of what is inside
    float t0 = 0, t1 = 1;
    if (fabsf(d.x) >= 1e-18f) { const float a = (-h.x - p0.x) / d.x, b = (h.x - p0.x) / d.x; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    if (fabsf(d.y) >= 1e-18f) { const float a = (-h.y - p0.y) / d.y, b = (h.y - p0.y) / d.y; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    if (fabsf(d.z) >= 1e-18f) { const float a = (-h.z - p0.z) / d.z, b = (h.z - p0.z) / d.z; t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    if (t1 >= t0) t = 0.5f * (t0 + t1);
    p = p0 + t * d;
    q = p;
    const float ex = h.x - fabsf(p.x), ey = h.y - fabsf(p.y), ez = h.z - fabsf(p.z);
    float best = ex;
    axis = 0;
    if (ey < best) { best = ey; axis = 1; }
    if (ez < best) { best = ez; axis = 2; }
    const float pc = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    sgn = pc >= 0 ? 1.0f : -1.0f;
    if (axis == 0) q.x = sgn * h.x; else if (axis == 1) q.y = sgn * h.y; else q.z = sgn * h.z;
    return -best;
}

// narrow phase of one primitive pair: surface points (world), unit normal from b to a, signed distance
template <bool BOX>
SC_DEV bool pair_contact(const float *S, const SelfTables &ST, int pr, float margin, V3 &pa, V3 &pb, V3 &n, float &dist) {
    const int a = ST.pair_a[pr], b = ST.pair_b[pr];
    if (BOX && ST.nbox > 0 && (ST.prim_box[a] >= 0 || ST.prim_box[b] >= 0)) {
        // a box is a box; when both are, the second one is stood in for by its capsule
        const bool a_is_box = ST.prim_box[a] >= 0;
        const int pbx = a_is_box ? a : b, psg = a_is_box ? b : a, body = ST.prim_body[pbx];
        const float *bx = ST.box + 15 * ST.prim_box[pbx], *core = ST.box_core + 7 * ST.prim_box[pbx];
        const M3 Rb = ldM(S, O_RW + 9 * body);
        const V3 s0 = prim_point(S, ST, psg, 0), s1 = prim_point(S, ST, psg, 1);
        const float rs = ST.prim[8 * psg + 6];
        {
            // most pairs that pass the bounding spheres are still centimetres apart (an arm hanging next to the torso): the
            // capsule that contains the box says so in closed form, and the exact routine below is only run for the rest
            const V3 o = ld3(S, O_PW + 3 * body);
            const V3 k0 = o + mul(Rb, mk(core[0], core[1], core[2])), k1 = o + mul(Rb, mk(core[3], core[4], core[5]));
            V3 c1, c2;
            seg_seg(k0, k1, s0, s1, c1, c2);
            const V3 dl = c1 - c2, ax = k1 - k0, rr = k0 - s0;
            // (seg_seg settles (nearly) parallel segments at an end point, which can overestimate their distance: the distance
            // of the two LINES never does)
            const float aa = dot(ax, ax), ra = dot(rr, ax);
            const float d2 = fminf(dot(dl, dl), fmaxf(0.0f, dot(rr, rr) - ra * ra / fmaxf(aa, 1e-18f)));
            const float lower = margin + fabsf(core[6]) + rs;
            if (d2 >= lower * lower) return false;
        }
        M3 Rw = Rb;
        if (core[6] < 0.0f) {               // (flag in the sign of the bound: the box is rotated in its body's frame)
            M3 Rl;
#pragma unroll
            for (int i = 0; i < 9; i++) Rl.a[i] = bx[3 + i];
            Rw = mul(Rb, Rl);
        }
        const V3 c = ld3(S, O_PW + 3 * body) + mul(Rb, mk(bx[0], bx[1], bx[2]));
        const V3 l0 = mulT(Rw, s0 - c), l1 = mulT(Rw, s1 - c);
        {
            // second bound, in the box's own frame: the segment's extent against the box grown by margin + radius, axis by axis
            const float g = margin + rs;
            if (fminf(l0.x, l1.x) > bx[12] + g || fmaxf(l0.x, l1.x) < -bx[12] - g || fminf(l0.y, l1.y) > bx[13] + g ||
                fmaxf(l0.y, l1.y) < -bx[13] - g || fminf(l0.z, l1.z) > bx[14] + g || fmaxf(l0.z, l1.z) < -bx[14] - g) return false;
        }
        V3 q, p;
        int axis;
        float sgn;
        const float len = box_segment(mk(bx[12], bx[13], bx[14]), l0, l1, q, p, axis, sgn);
        V3 nl;
        if (axis >= 0) nl = mk(axis == 0 ? sgn : 0.0f, axis == 1 ? sgn : 0.0f, axis == 2 ? sgn : 0.0f);
        else { if (len < 1e-9f) return false; nl = (1.0f / len) * (p - q); }
        dist = len - rs;
        const V3 nw = mul(Rw, nl);                                  // from the box to the swept segment
        const V3 pbox = c + mul(Rw, q), pseg = c + mul(Rw, p - rs * nl);
        n = a_is_box ? neg(nw) : nw;                                // from b to a
        pa = a_is_box ? pbox : pseg; pb = a_is_box ? pseg : pbox;
        return true;
    }

    V3 c1, c2;
    seg_seg(prim_point(S, ST, a, 0), prim_point(S, ST, a, 1), prim_point(S, ST, b, 0), prim_point(S, ST, b, 1), c1, c2);
    const V3 d = c1 - c2;
    const float len2 = dot(d, d);
    if (!(len2 > 1e-18f)) return false;
    const float len = sqrtf(len2), ra = ST.prim[8 * a + 6], rb = ST.prim[8 * b + 6];
    n = (1.0f / len) * d;
    dist = len - ra - rb;
    pa = c1 - ra * n; pb = c2 + rb * n;
    return true;
}

// u = unified candidate index: primitive pair u < npair, else ground point u - npair.  Link-vs-link contacts come first in
// the contact list: a column only evaluates the rows at or after its own contact, so the (few) two-body columns pay for
// the two-body rows and the ground columns never walk to a second body.
template <class TT>
SC_DEV void write_contact(float *S, const TT &T, const SelfTables &ST, int j, int u, float z) {
    if (u >= ST.npair) {
        const int c = u - ST.npair, b = T.cand_body[c];
        const float *pt = T.cand + 4 * c;
        F(O_CB + j) = i2f(b);
        const V3 o = mul(ldM(S, O_RW + 9 * b), mk(pt[0], pt[1], pt[2]));      // lowest point of the sphere: centre - r ez
        F(O_CR + 3 * j) = o.x;
        F(O_CR + 3 * j + 1) = o.y;
        F(O_CR + 3 * j + 2) = o.z - pt[3];
    } else {
        const int pr = u;
        const int ba = ST.prim_body[ST.pair_a[pr]], bb = ST.prim_body[ST.pair_b[pr]];
        V3 pa, pb, n;
        float d;
        pair_contact<TT::BOX>(S, ST, pr, 1.0e18f, pa, pb, n, d);
        F(O_CB + j) = i2f(ba | ((bb + 1) << 8));
        st3(S, O_CR + 3 * j, pa - ld3(S, O_PW + 3 * ba));
        st3(S, O_CR2 + 3 * j, pb - ld3(S, O_PW + 3 * bb));
        st3(S, O_CN + 3 * j, n);
    }
    F(O_CD + j) = z;
    F(O_CU + j) = i2f(u);
}

template <class TT>
SC_PHASE void collide(float *sm, const TT &T, const SelfTables &ST, int L0, int L1 SC_PROF_ARGS) {
    SC_STAGE {
        SC_LANE;
        const int c0 = slot * T.cand_per_lane;
        int hits = 0;
#pragma unroll
        for (int i = 0; i < CAND_PER_LANE_MAX; i++) {
            const int c = c0 + i;
            if (i < T.cand_per_lane && c < T.ncand) {
                const int b = T.cand_body[c];
                const float *pt = T.cand + 4 * c;
                const float z = F(O_PW + 3 * b + 2) + F(O_RW + 9 * b + 6) * pt[0] + F(O_RW + 9 * b + 7) * pt[1] +
                                F(O_RW + 9 * b + 8) * pt[2] - pt[3];
                F(O_ZC + ST.npair + c) = z;
                hits += (z < T.cthresh ? 1 : 0) + (z <= 0.0f ? 1 << 16 : 0);      // inside the margin | touching
            }
        }
        F(O_CNT + slot) = i2f(hits);
        if (ST.npair > 0)
#pragma unroll
            for (int i = 0; i < PRIMS_PER_LANE_MAX; i++) {
                const int p = slot * ST.prims_per_lane + i;
                if (i < ST.prims_per_lane && p < ST.nprim) {
                    const int b = ST.prim_body[p];
                    const float *q = ST.prim + 8 * p;
                    const V3 mid = mk(0.5f * (q[0] + q[3]), 0.5f * (q[1] + q[4]), 0.5f * (q[2] + q[5]));
                    st3(S, O_PCEN + 3 * p, ld3(S, O_PW + 3 * b) + mul(ldM(S, O_RW + 9 * b), mid));
                }
            }
    }
    SC_SYNC();
    SC_STAGE {
        SC_LANE;
        // broad phase: bounding spheres of the primitives; the rare survivors go through the narrow phase afterwards
        int hits = 0;
        unsigned near = 0;
#pragma unroll
        for (int i = 0; i < PAIRS_PER_LANE_MAX; i++) {
            const int pr = slot * ST.pairs_per_lane + i;
            if (i < ST.pairs_per_lane && pr < ST.npair) {
                const V3 d = ld3(S, O_PCEN + 3 * ST.pair_a[pr]) - ld3(S, O_PCEN + 3 * ST.pair_b[pr]);
                near |= (dot(d, d) < ST.pair_r2[pr] ? 1u : 0u) << i;
                F(O_ZC + pr) = NO_HIT;
            }
        }
        while (near) {
#if defined(__CUDA_ARCH__)
            const int i = __ffs(near) - 1;
#else
            const int i = __builtin_ctz(near);
#endif
            near &= near - 1;
            const int pr = slot * ST.pairs_per_lane + i;
            V3 pa, pb, n;
            float dist;
            if (pair_contact<TT::BOX>(S, ST, pr, T.cthresh, pa, pb, n, dist)) {
                F(O_ZC + pr) = dist;
                hits += (dist < T.cthresh ? 1 : 0) + (dist <= 0.0f ? 1 << 16 : 0);
            }
        }
        F(O_CNT + LPS + slot) = i2f(hits);
    }
    SC_SYNC();
    SC_STAGE {
        SC_LANE;
        int ofs = 0, ofs2 = 0, total = 0, touching = 0;   // ofs: this lane's first ground slot, ofs2: its first pair slot
#pragma unroll
        for (int l = 0; l < LPS; l++) {
            const int v = f2i(F(O_CNT + l)), v2 = f2i(F(O_CNT + LPS + l));
            const int n = v & 0xffff, n2 = v2 & 0xffff;
            ofs += (l < slot ? n : 0) + n2;
            ofs2 += l < slot ? n2 : 0;
            total += n + n2;
            touching += (v >> 16) + (v2 >> 16);
        }
#if defined(SC_PROFILE) && defined(__CUDACC__)
        if (total > T.maxc) sc_prof_acc[23]++;
        sc_prof_acc[24] += total < T.maxc ? total : T.maxc;
#endif
        const bool list = total > T.maxc && total <= HL_MAX;   // too many: compact (z, u) first, pick in parallel below
        if (total <= T.maxc || list) {
            if ((f2i(F(O_CNT + LPS + slot)) & 0xffff) > 0)
                for (int i = 0; i < ST.pairs_per_lane; i++) {
                    const int pr = slot * ST.pairs_per_lane + i;
                    if (pr < ST.npair) {
                        const float z = F(O_ZC + pr);
                        if (z < T.cthresh) {
                            if (list) { F(O_HL + 2 * ofs2) = z; F(O_HL + 2 * ofs2 + 1) = i2f(pr); ofs2++; }
                            else write_contact(S, T, ST, ofs2++, pr, z);
                        }
                    }
                }
            if ((f2i(F(O_CNT + slot)) & 0xffff) > 0)
                for (int i = 0; i < T.cand_per_lane; i++) {
                    const int c = slot * T.cand_per_lane + i;
                    if (c < T.ncand) {
                        const float z = F(O_ZC + ST.npair + c);
                        if (z < T.cthresh) {
                            if (list) { F(O_HL + 2 * ofs) = z; F(O_HL + 2 * ofs + 1) = i2f(ST.npair + c); ofs++; }
                            else write_contact(S, T, ST, ofs++, ST.npair + c, z);
                        }
                    }
                }
        }
        if (slot == 0) {
            F(O_NC) = i2f(total);
            // over the cap: the candidates dropped are the farthest ones -- speculative points that do not touch yet, unless
            // more than maxc points touch (counted separately, in the upper half)
            if (total > T.maxc) F(O_OVF) = i2f(f2i(F(O_OVF)) + 1 + (touching > T.maxc ? 1 << 16 : 0));
        }
    }
    SC_SYNC();
    if (tile_max_int(sm, O_NC) <= T.maxc) return;
    // ---- more than maxc candidates inside the margin somewhere in the tile: keep the maxc deepest (ties: lower index),
    // in candidate order.  Rank every hit against the compacted list; lists longer than HL_MAX fall back to lane 0.
    SC_STAGE {
        SC_LANE;
        const int total = f2i(F(O_NC));
        unsigned keep = 0;
        if (total > T.maxc && total <= HL_MAX) {
            for (int h = slot; h < total; h += LPS) {
                const float zh = F(O_HL + 2 * h);
                int rank = 0;
                for (int g = 0; g < total; g++) {
                    const float zg = F(O_HL + 2 * g);
                    rank += (zg < zh || (zg == zh && g < h)) ? 1 : 0;
                }
                keep |= (rank < T.maxc ? 1u : 0u) << h;
            }
        } else if (total > HL_MAX && slot == 0) {
            // O_CB holds candidate indices and O_CD distances while the deepest maxc are being chosen
            int n = 0;
            for (int c = 0; c < T.ncand + ST.npair; c++) {
                const float z = F(O_ZC + c);
                if (!(z < T.cthresh)) continue;
                if (n < T.maxc) { F(O_CB + n) = i2f(c); F(O_CD + n) = z; n++; continue; }
                int worst = 0;
                float zw = F(O_CD);
                for (int j = 1; j < n; j++) { const float zj = F(O_CD + j); if (zj >= zw) { worst = j; zw = zj; } }
                if (z < zw) {
                    for (int j = worst; j + 1 < n; j++) { F(O_CB + j) = F(O_CB + j + 1); F(O_CD + j) = F(O_CD + j + 1); }
                    F(O_CB + n - 1) = i2f(c); F(O_CD + n - 1) = z;
                }
            }
            for (int j = 0; j < n; j++) write_contact(S, T, ST, j, f2i(F(O_CB + j)), F(O_CD + j));
        }
        F(O_CNT + slot) = i2f((int)keep);
        if (slot == 0) F(O_CNT + LPS) = i2f(total);      // O_NC is rewritten in the next stage while others still need the count
    }
    SC_SYNC();
    SC_STAGE {
        SC_LANE;
        const int total = f2i(F(O_CNT + LPS));
        if (total > T.maxc && total <= HL_MAX) {
            unsigned keep = 0;
#pragma unroll
            for (int l = 0; l < LPS; l++) keep |= (unsigned)f2i(F(O_CNT + l));
            for (int h = slot; h < total; h += LPS)
                if ((keep >> h) & 1u) {
#if defined(__CUDA_ARCH__)
                    const int pos = __popc(keep & ((1u << h) - 1u));
#else
                    const int pos = __builtin_popcount(keep & ((1u << h) - 1u));
#endif
                    write_contact(S, T, ST, pos, f2i(F(O_HL + 2 * h + 1)), F(O_HL + 2 * h));
                }
        }
        if (slot == 0 && total > T.maxc) F(O_NC) = i2f(T.maxc);
    }
    SC_SYNC();
}

// Contact cache (every sub-step, after collide): contact j of the new list starts the sweep from warmstart x the normal
// impulse the same candidate ended the previous sub-step with (0 if it is new), and the cache is rewritten to describe the
// new list.  O_RHS is free at this point and carries the values over the barrier.
template <class TT>
SC_PHASE void contact_cache(float *sm, const TT &T, int L0, int L1) {
    static_assert(MAXC <= LPS * 2, "two contacts per lane at most");
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC)), pnc = f2i(F(O_PNC));
        for (int j = slot; j < MAXC; j += LPS) {
            float l0 = 0.0f;
            if (j < nc) {
                const int u = f2i(F(O_CU + j));
                for (int o = 0; o < pnc; o++) if (f2i(F(O_PCU + o)) == u) l0 = T.warmstart * F(O_PLAM + o);
            }
            F(O_RHS + j) = l0;
        }
    }
    SC_SYNC();
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int j = slot; j < MAXC; j += LPS) {
            F(O_PLAM + j) = F(O_RHS + j);
            F(O_PCU + j) = j < nc ? F(O_CU + j) : i2f(-1);
        }
        if (slot == 0) F(O_PNC) = i2f(nc);
    }
    SC_SYNC();
}

// One column of the contact-space matrix A = J M^-1 J^T, computed by one lane without barriers: the response of
// the articulated system (inertias left in shared memory by the dynamics ABA) to a unit impulse along direction k
// of contact j.  The impulse walks up the tree to the root (recording u = -S^T p per level in registers), the 6x6
// base block is solved with the Cholesky factor kept from the ABA, and accelerations walk back down to the bodies
// of the contacts i >= j; rows >= the column index land in the packed lower triangle O_AM.  A link-vs-link contact
// is the superposition of two such single-body problems (+f on body a, -f on body b): the second pass adds.
SC_CALL void contact_column(float *S, const Tables &T, int nc, int c, int i0) {
    const int j = c / 3, k = c - 3 * j;
    const int cbj = f2i(F(O_CB + j));
    const V3 f0 = contact_frame_of(S, j, cbj).d[k];
#pragma unroll 1
    for (int side = 0; side < 2; side++) {
        const int bj = side ? cb_b(cbj) : cb_a(cbj);
        if (bj < 0) break;
        const int dj = T.depth[bj];
        const V3 f = side ? neg(f0) : f0;
        V3 pn = neg(cross(ld3(S, (side ? O_CR2 : O_CR) + 3 * j), f)), pf = neg(f);
        V3 u[MAXLEV];
#pragma unroll
        for (int d = MAXLEV - 1; d >= 1; d--) {
            u[d] = mk(0, 0, 0);
            if (d <= dj) {
                const int b = T.anc[bj][d];
                // the contact phases see the dynamics pass's inertias (no armature): K A = 1, a spherical joint passes
                // no moment, so pn + A K u = pn + u = 0 and only the force part goes on to the parent
                u[d] = neg(pn);
                pf = pf + mul(ldM(S, O_UB + 9 * b), mul(ldS(S, O_DI + 6 * b), u[d]));
                pn = cross(ld3(S, O_RJ + 3 * b), pf);
            }
        }
        float L[21], x[6] = {-pn.x, -pn.y, -pn.z, -pf.x, -pf.y, -pf.z};
#pragma unroll
        for (int i = 0; i < 21; i++) L[i] = F(O_I0L + i);
        solve6(L, x);
        int bprev = -1;
        V3 awc = mk(0, 0, 0), alc = mk(0, 0, 0);
        // The walk over the contacts starts at i0 <= j, the first contact of the call's earliest column, for every lane:
        // the lanes of a sample then walk down the tree for a new body in the same iteration (one pass of the sweep code
        // for all of them) instead of each at its own j + t; rows before the column's own contact are not evaluated.
        // walk down the tree to body bi: angular and linear acceleration of its origin
#define SC_WALK_DOWN(bi, aw, al)                                                                                       \
        do {                                                                                                            \
            const int di = T.depth[bi];                                                                                 \
            aw = mk(x[0], x[1], x[2]); al = mk(x[3], x[4], x[5]);                                                       \
            _Pragma("unroll") for (int d = 1; d < MAXLEV; d++) {                                                        \
                if (d <= di) {                                                                                          \
                    const int b = T.anc[bi][d];                                                                         \
                    al = al + cross(aw, ld3(S, O_RJ + 3 * b));                                                          \
                    const V3 uu = (d <= dj && T.anc[bj][d] == b) ? u[d] : mk(0, 0, 0);                                  \
                    aw = mul(ldS(S, O_DI + 6 * b), uu - mulT(ldM(S, O_UB + 9 * b), al));   /* K (uu - B^T al), K A = 1 */ \
                }                                                                                                       \
            }                                                                                                           \
        } while (0)
        for (int i = i0; i < nc; i++) {
            const int cbi = f2i(F(O_CB + i)), ba = cb_a(cbi), bb = cb_b(cbi);
            if (i < j && (bb >= 0 || ba == bprev)) continue;
            // body a's response is remembered for the contacts that follow on the same body (remembering the last four
            // bodies instead was measured: bit-identical and 5 % slower per window -- the select chains cost more than the
            // walks they save); the ground-contact row, nearly all of them, is straight-line code after it
            if (ba != bprev) { SC_WALK_DOWN(ba, awc, alc); bprev = ba; }
            if (i < j) continue;              // only the body's response was wanted, for the rows to come
            V3 a = alc + cross(awc, ld3(S, O_CR + 3 * i));
            if (bb >= 0) {                    // link-link: minus body b's acceleration at its contact point
                V3 aw, al;
                SC_WALK_DOWN(bb, aw, al);
                a = a - (al + cross(aw, ld3(S, O_CR2 + 3 * i)));
            }
#undef SC_WALK_DOWN
            // projections on contact i's frame; a ground contact's frame is +z, -y, +x: components, no arithmetic
            V3 pr = mk(a.z, -a.y, a.x);
            if (cbi >> 8) {
                const Frame fi = contact_frame(ld3(S, O_CN + 3 * i));
                pr = mk(dot(fi.d[0], a), dot(fi.d[1], a), dot(fi.d[2], a));
            }
            const int r0 = 3 * i, o0 = O_AM + r0 * (r0 + 1) / 2 + c;
            if (side) {
                if (r0 >= c) F(o0) += pr.x;
                if (r0 + 1 >= c) F(o0 + r0 + 1) += pr.y;
                if (r0 + 2 >= c) F(o0 + 2 * r0 + 3) += pr.z;
            } else {
                if (r0 >= c) F(o0) = pr.x;
                if (r0 + 1 >= c) F(o0 + r0 + 1) = pr.y;
                if (r0 + 2 >= c) F(o0 + 2 * r0 + 3) = pr.z;
            }
        }
    }
}

#if defined(__CUDACC__)
// Device version of the Gauss-Seidel sweep.  The sweep visits the rows in Bullet's order (all normals, then the friction
// pairs): position s = 0..3*NCT-1 is row gs_row(s).  Lane `slot` of a sample owns the three consecutive positions
// 3*slot .. 3*slot+2; those rows of the (symmetric) matrix, their residuals and impulses stay in registers for all
// iterations, the owner of the updated row broadcasts the impulse change to the sample's lanes with one shuffle, and
// every lane folds it into its residuals.  Inside a lane's run of three the next row does not wait for that shuffle:
// the owner already knows its own change, so the residual its next row needs is formed from the local value
// (identical bits), and the broadcast only has to arrive before the next lane's run.  Rows beyond 3*nc are zero rows
// and never move.  The sweep is unrolled for NCT contacts, so the control flow is uniform and branch free.
__host__ __device__ constexpr int gs_row(int s, int nct) { return s < nct ? 3 * s : 3 * ((s - nct) / 2) + 1 + ((s - nct) & 1); }

template <int NCT, class TT>
__device__ __forceinline__ void gs_registers(float *sm, const TT &T) {
    const int lane = threadIdx.x & 31;
    float *S = sm + (lane & (TILE - 1));
    const int slot = lane / TILE;
    const int nr = 3 * f2i(F(O_NC));
    constexpr int NRT = 3 * NCT, RPL = 3;                           // rows per lane: one run of the sweep
    static_assert(NRT == RPL * LPS, "one run of three sweep positions per lane");
    float a[RPL][NRT], res[RPL], lam[RPL], dgi[RPL], lamn[NCT], muc[NCT];
#pragma unroll
    for (int m = 0; m < RPL; m++) {
        const int row = gs_row(RPL * slot + m, NCT);
        const bool live = row < nr;
        // element (row, c) of the packed lower triangle: in the row's own run for c <= row, else in row c at a compile-time
        // offset.  One address select and one unconditional load per element -- entries beyond the live system are whatever
        // the scratch block holds and are replaced by zeros, never used (r1 computed max / min / triangular index per element
        // behind a branch: 660 instructions and 35 branches before the first sweep, half of the Gauss-Seidel phase's time)
        const float *rowp = S + (O_AM + row * (row + 1) / 2) * TILE, *colp = S + (O_AM + row) * TILE;
#pragma unroll
        for (int c = 0; c < NRT; c++) {
            const float *p = row >= c ? rowp + c * TILE : colp + (c * (c + 1) / 2) * TILE;
            const float v = *p;
            a[m][c] = (live && c < nr) ? v : 0.0f;
        }
        res[m] = live ? F(O_RHS + row) : 0.0f;
        dgi[m] = live ? 1.0f / rowp[row * TILE] : 0.0f;
    }
    // warm start: the normal rows begin at the cached impulses (contact_cache), friction rows at zero
#pragma unroll
    for (int c = 0; c < NCT; c++) {
        lamn[c] = 3 * c < nr ? F(O_PLAM + c) : 0.0f;
        muc[c] = (f2i(F(O_CB + c)) >> 8) ? T.mu_self : T.mu;
#pragma unroll
        for (int m = 0; m < RPL; m++) res[m] -= a[m][3 * c] * lamn[c];
    }
#pragma unroll
    for (int m = 0; m < RPL; m++) {
        const int row = gs_row(RPL * slot + m, NCT);
        lam[m] = 0.0f;
#pragma unroll
        for (int c = 0; c < NCT; c++) if (row == 3 * c) lam[m] = lamn[c];
    }
    const int src0 = lane & (TILE - 1);
    for (int it = 0; it < T.niter; it++) {
        float spec = 0.0f;
#pragma unroll
        for (int s = 0; s < NRT; s++) {
            const int r = gs_row(s, NCT), c = r / 3, k = r - 3 * c, owner = s / RPL, m = s % RPL;
            // the owner's residual of this row: inside a run, the value formed from its own previous change
            const float rcur = m > 0 ? spec : res[m];
            float nl = lam[m] + rcur * dgi[m];
            if (k == 0) nl = fmaxf(0.0f, nl);
            else { const float lim = muc[c] * lamn[c]; nl = fminf(lim, fmaxf(-lim, nl)); }
            // only the owner's value is read (by the shuffle below and by its own next row): the other lanes' is never used
            const float dloc = nl - lam[m];
            if (slot == owner) lam[m] = nl;
            if (m + 1 < RPL) spec = res[m + 1] - a[m + 1][r] * dloc;
            const float d = __shfl_sync(0xffffffffu, dloc, src0 + TILE * owner);
            if (k == 0) lamn[c] += d;
#pragma unroll
            for (int mm = 0; mm < RPL; mm++) res[mm] -= a[mm][r] * d;
        }
    }
#pragma unroll
    for (int m = 0; m < RPL; m++) {
        const int row = gs_row(RPL * slot + m, NCT);
        if (row < nr) {
            F(O_LAM + row) = lam[m];          // same words as the right-hand sides this lane read above
            if (row % 3 == 0) F(O_PLAM + row / 3) = lam[m];
        }
    }
    __syncwarp();
}
#endif

// The assembly is dealt in "column calls": in call m every lane of the warp computes column asm_column(m, slot) of its
// sample (dealing the columns of all four samples over all 32 lanes was measured 13 % slower: lanes of one sample share 8
// of the 32 banks, the tile's interleave only keeps different samples apart).  Columns slot, 15 - slot, 16 + slot: a
// column evaluates the rows at or after its own, so the calls get cheaper (8 contacts: 8, 6 and 3 contact rows).
constexpr int ASM_CALLS = (3 * MAXC + LPS - 1) / LPS;
SC_DEV int asm_column(int m, int slot) { return m == 1 ? 2 * LPS - 1 - slot : slot + LPS * m; }
SC_DEV int asm_first_contact(int m) { return (LPS * m) / 3; }      // contact of the call's lowest column
SC_DEV int asm_calls(int ncmax) {       // calls a tile with at most ncmax contacts per sample needs
    int n = 0;
    for (int m = 0; m < ASM_CALLS; m++) n += 3 * ncmax > (m == 1 ? LPS : LPS * m);
    return n;
}

constexpr int COOP_WORDS = 64;          // per CTA, behind the tiles: the pool's bookkeeping (assembly_shared)
#if defined(__CUDACC__) && SC_SHARE_ASSEMBLY
// Shared assembly.  The column calls of a CTA's tiles are a common pool: every warp publishes its tile (number of calls,
// claim counter, "ready" = the epoch of this sub-step) once the tile's contact list and right-hand sides are written,
// works through its own calls, and then takes unclaimed calls of the other tiles until every tile of the CTA is claimed
// out.  A column only reads its tile (inertias, contact list) and writes its own entries of the packed matrix, so who
// computes it does not change a bit of the result; the CTA barrier that follows the assembly makes the helpers' writes
// visible to the owner.  Measured reason: a tile needs one to three calls of about 10 / 8.5 / 5.5 k cycles depending on
// its samples' contact counts, link-link columns cost double, and 10.7 % of all warp time was spent at that barrier
// waiting for the slowest tile (profiles/r2zz_final_phase_stalls.md; 4-5 % with the pool, profiles/r2_final3_phase_stalls.md).
// coop: [0,16) ready epoch, [16,32) next call, [32,48) number of calls, per warp; [48,56) the profile build's maxima.
template <class TT>
__device__ __forceinline__ void assembly_shared(float *sm, const TT &T, int ncmax, int *coop_, int epoch SC_PROF_ARGS) {
    volatile int *coop = coop_;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#if defined(SC_PROFILE)
    // developer build: when (cycles after the barrier in front of collide) this warp's tile is ready, its own calls are
    // done, and it leaves the pool -- mean over the warps in [32..34], the CTA's maximum per sub-step in [35..37]
    const bool rec = sc_prof_acc[38] >= 0;                    // [38]: cycles since that barrier at the call, < 0: a warp without a tile
    const long long t0 = clock64() - sc_prof_acc[38];
    int *pmax = coop_ + 48 + 4 * (epoch & 1);
    bool own_done = !rec;
    if (rec) { const int dt = (int)(clock64() - t0); sc_prof_acc[32] += dt; if (lane == 0) atomicMax(pmax, dt); }
#endif
    if (lane == 0) {
        coop[32 + warp] = asm_calls(ncmax);
        coop[16 + warp] = 0;
        __threadfence_block();
        coop[warp] = epoch;
    }
    __syncwarp();
    unsigned pending = (1u << nw) - 1u;
    int t = warp, idle = 0;
    while (pending) {
        int m = -1, exhausted = 0;
        if (lane == 0 && coop[t] == epoch) {
            __threadfence_block();
            const int n = coop[32 + t];
            if (coop[16 + t] < n) m = atomicAdd(coop_ + 16 + t, 1);
            if (m < 0 || m >= n) { m = -1; exhausted = 1; }
        }
        m = __shfl_sync(0xffffffffu, m, 0);
        exhausted = __shfl_sync(0xffffffffu, exhausted, 0);
        if (m >= 0) {
            float *S = sm + (t - warp) * WORDS_PER_TILE + (lane & (TILE - 1));
            const int slot = lane / TILE, nc = f2i(F(O_NC)), c = asm_column(m, slot);
#if defined(SC_PROFILE)
            const long long tc0 = clock64();
#endif
            if (c < 3 * nc) contact_column(S, T, nc, c, asm_first_contact(m));
            __syncwarp();
#if defined(SC_PROFILE)
            if (rec) { sc_prof_acc[t == warp ? 25 : 31] += clock64() - tc0; sc_prof_acc[14] += t == warp ? 1 : 0x1000000; }
            // trace of CTA 0: (epoch, warp, tile, call, two-body tile?, start, end) of every call of a few sub-steps
            if (sc_trace && blockIdx.x == 0 && epoch >= 150 && epoch < 156 && lane == 0) {
                const int k = (int)atomicAdd((unsigned long long *)sc_trace, 1ull);
                if (k < 400) {
                    long long *e = sc_trace + 8 + 4 * k;
                    e[0] = epoch;
                    e[1] = warp | t << 8 | m << 16 | (f2i(sm[(t - warp) * WORDS_PER_TILE + O_TM * TILE]) >= TM_TWO_BODY ? 1 << 24 : 0);
                    e[2] = tc0 - t0; e[3] = clock64() - t0;
                }
            }
#endif
            idle = 0;
            continue;                                   // the same tile again: its next call may still be free
        }
#if defined(SC_PROFILE)
        if (t == warp && exhausted && !own_done) {
            own_done = true;
            const int dt = (int)(clock64() - t0); sc_prof_acc[33] += dt; if (lane == 0) atomicMax(pmax + 1, dt);
        }
#endif
        if (exhausted) pending &= ~(1u << t);
        else if (++idle >= nw) { __nanosleep(SC_POLL_NS); idle = 0; }       // a full round over tiles that are not ready yet
        if (pending) {                                  // next tile that is not known to be claimed out
            const unsigned above = pending & ~((2u << t) - 1u);
            t = __ffs(above ? above : pending) - 1;
        }
    }
#if defined(SC_PROFILE)
    if (rec) { const int dt = (int)(clock64() - t0); sc_prof_acc[34] += dt; if (lane == 0) atomicMax(pmax + 2, dt); }
#endif
}
#endif

template <class TT>
SC_PHASE void contact_solve(float *sm, const TT &T, int L0, int L1, int ncmax, const int part SC_PROF_ARGS) {
    if (part == 0) {
    // ---- right-hand sides from the predicted velocities; mask of the bodies on a contact's root path ------------
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int j = slot; j < nc; j += LPS) {
            const int cb = f2i(F(O_CB + j)), ba = cb_a(cb), bb = cb_b(cb);
            V3 vp = ld3(S, O_TW + 6 * ba + 3) + cross(ld3(S, O_TW + 6 * ba), ld3(S, O_CR + 3 * j));
            if (bb >= 0) vp = vp - ld3(S, O_TW + 6 * bb + 3) - cross(ld3(S, O_TW + 6 * bb), ld3(S, O_CR2 + 3 * j));
            const Frame fr = contact_frame_of(S, j, cb);
            const float dist = F(O_CD + j);
            F(O_RHS + 3 * j) = -dot(fr.d[0], vp) - (dist > 0 ? dist / T.h : dist * T.erp / T.h);
            F(O_RHS + 3 * j + 1) = -dot(fr.d[1], vp);
            F(O_RHS + 3 * j + 2) = -dot(fr.d[2], vp);
        }
        if (slot == LPS - 1) {
            int mask = 0, direct = 0;
            for (int j = 0; j < nc; j++) {
                const int cb = f2i(F(O_CB + j));
                mask |= T.anc_mask[cb_a(cb)] | (cb_b(cb) >= 0 ? T.anc_mask[cb_b(cb)] | TM_TWO_BODY : 0);
                direct |= 1 << cb_a(cb) | (cb_b(cb) >= 0 ? 1 << cb_b(cb) : 0);
            }
            F(O_TM) = i2f(mask);
            F(O_DM) = i2f(direct);
        }
#if !defined(__CUDACC__)
        if (slot == 0) { F(O_DEL) = 0; F(O_DEL + 1) = 0; }
#endif
    }
    SC_SYNC();
    SC_T(7);
    }
    if (part == 3) {
    // ---- contact-space matrix: the lanes of a sample take its columns round-robin, no barriers in between --------
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int m = 0; m < ASM_CALLS; m++) {
            const int c = asm_column(m, slot);
            if (c < 3 * nc) contact_column(S, T, nc, c, asm_first_contact(m));
        }
    }
    SC_SYNC();
    SC_T(8);
    }
    if (part == 1) {
    // ---- projected Gauss-Seidel, Bullet row order: all normals, then the friction pairs -------------------------
#if defined(__CUDACC__)
    // one sweep size for every warp: variants unrolled for 2 / 4 / 6 contacts do less work for tiles with few contacts,
    // but put the CTA's warps into different code at the same time, and instruction fetch is what this kernel is short
    // of (measured: 1 variant 32.4 ms, 4 variants 32.7 ms per walk window; skipping the rows of absent contacts with
    // warp-uniform branches inside the one sweep: 30.0 vs 29.2 ms -- the straight-line sweep schedules better)
    (void)ncmax;
    gs_registers<MAXC>(sm, T);
    SC_T(9);
#else
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int r = slot; r < 3 * nc; r += LPS) {
            F(O_DGI + r) = 1.0f / F(O_AM + r * (r + 1) / 2 + r);
            float res = F(O_RHS + r);
            for (int c = 0; c < nc; c++) {
                const int cc = 3 * c, hi = r > cc ? r : cc, lo = r > cc ? cc : r;
                res -= F(O_AM + hi * (hi + 1) / 2 + lo) * F(O_PLAM + c);
            }
            F(O_RES + r) = res;
        }
    }
    SC_SYNC();
    SC_STAGE {      // O_LAM shares its words with O_RHS: start values go in once the residuals are formed
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int r = slot; r < 3 * nc; r += LPS) F(O_LAM + r) = (r % 3 == 0) ? F(O_PLAM + r / 3) : 0.0f;
    }
    SC_SYNC();
    const int per_iter = 3 * ncmax, total = T.niter * per_iter;
    for (int t = 0; t < total; t++) {
        SC_STAGE {
            SC_LANE;
            const int nc = f2i(F(O_NC));
            // apply the impulse change decided in the previous stage to the residuals this lane owns
            int rp = -1;
            if (t > 0) {
                const int tl = (t - 1) % per_iter;
                const int c = tl < ncmax ? tl : (tl - ncmax) >> 1;
                if (c < nc) rp = tl < ncmax ? 3 * c : 3 * c + 1 + ((tl - ncmax) & 1);
            }
            const float dl = F(O_DEL + ((t + 1) & 1));
            if (rp >= 0 && dl != 0.0f)
                for (int i = slot; i < 3 * nc; i += LPS) {
                    const int hi = i > rp ? i : rp, lo = i > rp ? rp : i;
                    F(O_RES + i) -= F(O_AM + hi * (hi + 1) / 2 + lo) * dl;
                }
            // the owner of the next row decides its impulse change
            {
                const int tl = t % per_iter;
                const int c = tl < ncmax ? tl : (tl - ncmax) >> 1;
                if (c < nc) {
                    const int k = tl < ncmax ? 0 : 1 + ((tl - ncmax) & 1);
                    const int r = 3 * c + k;
                    if ((r % LPS) == slot) {
                        const float lam = F(O_LAM + r);
                        float nl = lam + F(O_RES + r) * F(O_DGI + r);
                        if (k == 0) nl = fmaxf(nl, 0.0f);
                        else { const float lim = ((f2i(F(O_CB + c)) >> 8) ? T.mu_self : T.mu) * F(O_LAM + 3 * c); nl = fminf(lim, fmaxf(-lim, nl)); }
                        F(O_LAM + r) = nl;
                        F(O_DEL + (t & 1)) = nl - lam;
                    }
                } else if (slot == 0) F(O_DEL + (t & 1)) = 0.0f;
            }
        }
        SC_SYNC();
    }
    SC_STAGE {
        SC_LANE;
        const int nc = f2i(F(O_NC));
        for (int c = slot; c < nc; c += LPS) F(O_PLAM + c) = F(O_LAM + 3 * c);
    }
    SC_SYNC();
#endif
    }
    if (part != 2) return;
    // ---- apply the impulses: one articulated response over the tree ------------------------------------------
    for (int d = T.nlev - 1; d >= 0; d--) {
        SC_STAGE {
            SC_LANE;
            const int b = T.level_start[d] + slot;
            const int mask = f2i(F(O_TM)), nc = f2i(F(O_NC));
            if (b < T.level_start[d + 1] && nc > 0 && ((mask >> b) & 1)) {
                V3 pn = mk(0, 0, 0), pf = mk(0, 0, 0);
                for (int k = 0; k < T.nchild[b]; k++) {
                    const int c = T.child[b][k];
                    if ((mask >> c) & 1) { pn = pn + ld3(S, O_PC + 6 * c); pf = pf + ld3(S, O_PC + 6 * c + 3); }
                }
                // (only a body that carries a contact itself looks through the list: most bodies on a root path do not)
                if ((f2i(F(O_DM)) >> b) & 1)
                for (int j = 0; j < nc; j++) {
                    const int cb = f2i(F(O_CB + j));
                    const bool ona = cb_a(cb) == b, onb = cb_b(cb) == b;
                    if (ona || onb) {
                        const Frame fr = contact_frame_of(S, j, cb);
                        const V3 f = F(O_LAM + 3 * j) * fr.d[0] + F(O_LAM + 3 * j + 1) * fr.d[1] + F(O_LAM + 3 * j + 2) * fr.d[2];
                        if (ona) { pn = pn - cross(ld3(S, O_CR + 3 * j), f); pf = pf - f; }
                        else { pn = pn + cross(ld3(S, O_CR2 + 3 * j), f); pf = pf + f; }
                    }
                }
                if (b == 0) {
                    float L[21], x[6] = {-pn.x, -pn.y, -pn.z, -pf.x, -pf.y, -pf.z};
#pragma unroll
                    for (int i = 0; i < 21; i++) L[i] = F(O_I0L + i);
                    solve6(L, x);
#pragma unroll
                    for (int i = 0; i < 6; i++) F(O_PC + i) = x[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) F(O_VB + i) += x[i];
                } else {
                    const V3 u = neg(pn);
                    st3(S, O_UU + 3 * b, u);
                    pf = pf + mul(ldM(S, O_UB + 9 * b), mul(ldS(S, O_DI + 6 * b), u));            // pn + A K u = 0 (K A = 1)
                    st3(S, O_PC + 6 * b, cross(ld3(S, O_RJ + 3 * b), pf));
                    st3(S, O_PC + 6 * b + 3, pf);
                }
            }
        }
        SC_SYNC();
    }
    for (int d = 1; d < T.nlev; d++) {
        SC_STAGE {
            SC_LANE;
            const int b = T.level_start[d] + slot;
            const int mask = f2i(F(O_TM)), nc = f2i(F(O_NC));
            if (b < T.level_start[d + 1] && nc > 0) {
                const int p = T.parent[b];
                const V3 aw = ld3(S, O_PC + 6 * p), al = ld3(S, O_PC + 6 * p + 3) + cross(aw, ld3(S, O_RJ + 3 * b));
                V3 u = mk(0, 0, 0);
                if ((mask >> b) & 1) u = ld3(S, O_UU + 3 * b);
                const V3 awn = mul(ldS(S, O_DI + 6 * b), u - mulT(ldM(S, O_UB + 9 * b), al));    // aw + K (u - A aw - B^T al), K A = 1
                st3(S, O_PC + 6 * b, awn);
                st3(S, O_PC + 6 * b + 3, al);
                st3(S, O_W + 3 * b, ld3(S, O_W + 3 * b) + mulT(ldM(S, O_RW + 9 * b), awn - aw));
            }
        }
        SC_SYNC();
    }
}

// ============================================================================================================
// integration (semi-implicit Euler, exponential-map quaternion update, Bullet's velocity clamp)
// ============================================================================================================
SC_DEV void quat_step(float *S, int o, V3 w, float h, bool left) {
    // exp(h w / 2): the half angle x = |w| h / 2 is at most 0.5 * sqrt(3) * maxvel * h (0.043 at the reference settings),
    // where the degree-7/6 Taylor polynomials are exact to float rounding (next terms x^8/9! and x^8/8! relative)
    const float x2 = 0.25f * h * h * dot(w, w);
    const float s = 0.5f * h * (1.0f + x2 * (-1.0f / 6.0f + x2 * (1.0f / 120.0f - x2 * (1.0f / 5040.0f))));     // sin(x) / |w|
    const float ew = 1.0f + x2 * (-0.5f + x2 * (1.0f / 24.0f - x2 * (1.0f / 720.0f)));                           // cos(x)
    const float ex = w.x * s, ey = w.y * s, ez = w.z * s;
    const float qx = F(o), qy = F(o + 1), qz = F(o + 2), qw = F(o + 3);
    float x, y, z, ww;
    if (left) {   // e (x) q
        x = ew * qx + ex * qw + ey * qz - ez * qy; y = ew * qy - ex * qz + ey * qw + ez * qx;
        z = ew * qz + ex * qy - ey * qx + ez * qw; ww = ew * qw - ex * qx - ey * qy - ez * qz;
    } else {      // q (x) e
        x = qw * ex + qx * ew + qy * ez - qz * ey; y = qw * ey - qx * ez + qy * ew + qz * ex;
        z = qw * ez + qx * ey - qy * ex + qz * ew; ww = qw * ew - qx * ex - qy * ey - qz * ez;
    }
    const float n = rsq(x * x + y * y + z * z + ww * ww);
    F(o) = x * n; F(o + 1) = y * n; F(o + 2) = z * n; F(o + 3) = ww * n;
}

template <class TT>
SC_PHASE void integrate(float *sm, const TT &T, int L0, int L1) {
    SC_STAGE {
        SC_LANE;
        const float mv = T.maxvel, h = T.h;
        for (int b = slot; b < T.nb; b += LPS) {
            if (b == 0) {
#pragma unroll
                for (int i = 0; i < 6; i++) F(O_VB + i) = fminf(mv, fmaxf(-mv, F(O_VB + i)));
                F(O_P0) += h * F(O_VB + 3); F(O_P0 + 1) += h * F(O_VB + 4); F(O_P0 + 2) += h * F(O_VB + 5);
                quat_step(S, O_Q, ld3(S, O_VB), h, true);
            } else {
                V3 w = ld3(S, O_W + 3 * b);
                w = mk(fminf(mv, fmaxf(-mv, w.x)), fminf(mv, fmaxf(-mv, w.y)), fminf(mv, fmaxf(-mv, w.z)));
                st3(S, O_W + 3 * b, w);
                quat_step(S, O_Q + 4 * b, w, h, false);
            }
        }
    }
    SC_SYNC();
}

// ============================================================================================================
// state I/O in the reference's 132-float layout (humanoid_stable_pd.py:137-165)
// ============================================================================================================
SC_DEV int state_word(const Tables &T, int i) {
    const int nj = T.nj;
    if (i < 3) return O_P0 + i;
    if (i < 7) return O_Q + (i - 3);
    if (i < 7 + 4 * nj) { const int j = (i - 7) >> 2; return O_Q + 4 * T.jbody[j] + ((i - 7) & 3); }
    i -= 7 + 4 * nj;
    if (i < 3) return O_VB + 3 + i;
    if (i < 6) return O_VB + (i - 3);
    i -= 6;
    return O_W + 3 * T.jbody[i / 3] + i % 3;
}

template <class TT>
SC_DEV void load_state(float *sm, const TT &T, int L0, int L1, const float *rows, const float *cache_rows, const int *parent_idx,
                       int children, int tile0, int S_total) {
    SC_STAGE {
        SC_LANE;
        int k = tile0 + (lane & (TILE - 1));
        if (k >= S_total) k = S_total - 1;   // padding lanes replay the last sample and never write
        const size_t pr = (size_t)(parent_idx ? parent_idx[k] : k / children);
        const float *src = rows + pr * SC_STATE_DIM;
        for (int i = slot; i < SC_STATE_DIM; i += LPS) F(state_word(T, i)) = src[i];
        // contact cache row: MAXC candidate ids (as floats, -1 = empty, filled from the front), MAXC normal impulses
        const float *cr = cache_rows ? cache_rows + pr * SC_CACHE_DIM : nullptr;
        for (int j = slot; j < MAXC; j += LPS) {
            F(O_PCU + j) = i2f(cr ? (int)cr[j] : -1);
            F(O_PLAM + j) = cr ? cr[MAXC + j] : 0.0f;
        }
        if (slot == 0) {
            int n = 0;
            if (cr) for (int j = 0; j < MAXC; j++) n += cr[j] >= 0.0f ? 1 : 0;
            F(O_PNC) = i2f(n);
            F(O_OVF) = i2f(0);
        }
    }
    SC_SYNC();
    SC_STAGE {
        SC_LANE;
        for (int b = slot; b < T.nb; b += LPS) {
            const int o = O_Q + 4 * b;
            // a quaternion this kernel wrote comes back unit to rounding and is left alone, so that resuming from
            // (state, cache) reproduces the uncut run bit for bit; anything else (host data) is normalised
            const float n2 = F(o) * F(o) + F(o + 1) * F(o + 1) + F(o + 2) * F(o + 2) + F(o + 3) * F(o + 3);
            if (fabsf(n2 - 1.0f) > 1e-6f) {
                const float n = rsq(n2);
                F(o) *= n; F(o + 1) *= n; F(o + 2) *= n; F(o + 3) *= n;
            }
        }
    }
    SC_SYNC();
}

// ============================================================================================================
// cost features: link CoM positions / velocities -> whole-body CoM, end-effector positions
// (humanoid_stable_pd.py:167-187, 218-222)
// ============================================================================================================
template <class TT>
SC_PHASE void features(float *sm, const TT &T, int L0, int L1) {
    fk_sweep(sm, T, L0, L1, true);
    SC_STAGE {
        SC_LANE;
        if (slot == 0) {
            V3 cp = mk(0, 0, 0), cv = mk(0, 0, 0);
            for (int i = 0; i < T.ncp; i++) {
                const int b = T.cp_body[i];
                const V3 off = mul(ldM(S, O_RW + 9 * b), tv3(T.cp_off, i));
                const V3 p = ld3(S, O_PW + 3 * b) + off;
                const V3 v = ld3(S, O_TW + 6 * b + 3) + cross(ld3(S, O_TW + 6 * b), off);
                cp = cp + T.cp_mass[i] * p;
                cv = cv + T.cp_mass[i] * v;
                for (int e = 0; e < T.nee; e++) if (T.ee[e] == i) st3(S, O_FEAT + 6 + 3 * e, p);
            }
            st3(S, O_FEAT, T.inv_total_mass * cp);
            st3(S, O_FEAT + 3, T.inv_total_mass * cv);
        }
    }
    SC_SYNC();
}

SC_DEV float quat_angle(float ax, float ay, float az, float aw, float bx, float by, float bz, float bw) {
    // rotation angle between two orientations, 2*acos(|<a,b>|) written through the relative quaternion's vector part
    const float dx = bw * ax - bx * aw - by * az + bz * ay;
    const float dy = bw * ay + bx * az - by * aw - bz * ax;
    const float dz = bw * az - bx * ay + by * ax - bz * aw;
    const float dw = bw * aw + bx * ax + by * ay + bz * az;
    return 2.0f * atan2f(sqrtf(dx * dx + dy * dy + dz * dz), fabsf(dw));
}

// compute_total_cost (humanoid_tracking_env.py:78-176); kin = target state row, kf = its 18 features
SC_DEV void cost_terms(const float *S, const Tables &T, const float *kin, const float *kf, float *out6) {
    const int nj = T.nj;
    float pose = 0;
    for (int b = 1; b < T.nb; b++) {
        const int j = T.jidx[b];
        const float *q = kin + 7 + 4 * j, *w = kin + 13 + 4 * nj + 3 * j;
        const float ang = quat_angle(F(O_Q + 4 * b), F(O_Q + 4 * b + 1), F(O_Q + 4 * b + 2), F(O_Q + 4 * b + 3), q[0], q[1], q[2], q[3]);
        const V3 dw = ld3(S, O_W + 3 * b) - mk(w[0], w[1], w[2]);
        pose += T.jw[b] * (ang * ang + 0.1f * dot(dw, dw));
    }
    pose /= (float)nj;
    const float ang = quat_angle(F(O_Q), F(O_Q + 1), F(O_Q + 2), F(O_Q + 3), kin[3], kin[4], kin[5], kin[6]);
    const float *kw = kin + 10 + 4 * nj;
    const V3 dw = ld3(S, O_VB) - mk(kw[0], kw[1], kw[2]);
    const float root = ang * ang + 0.1f * dot(dw, dw);
    float ee = 0, bal = 0;
    const V3 cs = ld3(S, O_FEAT), ck = mk(kf[0], kf[1], kf[2]);
    for (int e = 0; e < T.nee; e++) {
        const V3 ps = ld3(S, O_FEAT + 6 + 3 * e), pk = mk(kf[6 + 3 * e], kf[7 + 3 * e], kf[8 + 3 * e]);
        const float dz = ps.z - pk.z;
        ee += dz * dz;
        const float bx = (cs.x - ps.x) - (ck.x - pk.x), by = (cs.y - ps.y) - (ck.y - pk.y);
        bal += bx * bx + by * by;
    }
    ee /= (float)T.nee;
    bal /= (float)T.nee * T.height;
    const V3 dc = cs - ck, dcv = ld3(S, O_FEAT + 3) - mk(kf[3], kf[4], kf[5]);
    const float com = dot(dc, dc) + 0.1f * dot(dcv, dcv);
    out6[1] = pose; out6[2] = root; out6[3] = ee; out6[4] = bal; out6[5] = com;
    out6[0] = T.cw[0] * pose + T.cw[1] * root + T.cw[2] * ee + T.cw[3] * bal + T.cw[4] * com;
}

// ============================================================================================================
// one tile = 4 samples through a whole SamCon window
// ============================================================================================================
struct RolloutArgs {
    const float *parents;      // [E,132]
    const float *parent_cache; // [E,SC_CACHE_DIM] contact caches of the parents, or null (empty caches)
    const int *parent_idx;     // [S] or null
    int children;
    const float *actions;      // [S,4*nj]
    const float *targets;      // [nseq,132]
    const float *tfeat;        // [nseq,18]
    const int *sample_seq;     // [S] or null
    int S, nsteps;
    float *out_state, *out_cost, *out_info;   // any may be null
    float *out_feat;           // [S,18] or null (feature-only mode)
    float *out_cost6;          // [S,6] total + 5 terms, or null (sc_cost)
    float *out_cache;          // [S,SC_CACHE_DIM] or null
    int *out_overflow;         // [S] or null: sub-steps of the window in which the sample had more than max_contacts candidate
                               // points inside the margin (low 16 bits) / touching, distance <= 0 (high 16 bits)
    unsigned long long *stats; // [3] or null: sample-sub-steps with more candidates in the margin than max_contacts, sample-sub-steps
                               // simulated, sample-sub-steps with more TOUCHING candidates than max_contacts
    long long *prof;           // developer builds (-DSC_PROFILE): per-phase cycle counters, else null
    int tile_warps;            // set by the launcher: warps of a CTA that own a tile (the rest only help with the assembly)
};

template <class TT, int BM = 31>
SC_DEV void rollout_tile(float *sm, const TT &T, const RolloutArgs &a, int tile0, int L0, int L1, int *coop, int epoch0 SC_PROF_ARGS) {
    // coop / epoch0 (device only): the CTA's assembly pool and the number of sub-steps its earlier rounds have used it for
    (void)coop; (void)epoch0;
    SC_T(15);
    if (tile0 >= a.S) {
        // a warp without a tile in this round keeps the CTA's phase barriers company and helps with the others' assembly
        for (int step = 0; step < a.nsteps; step++) {
            SC_BARL(1); SC_BAR(256);
            for (int sub = 0; sub < T.nsub; sub++) {
                SC_BARL(2); SC_BAR(256); SC_BARL(4); SC_BARL(64);
#if defined(__CUDACC__) && SC_SHARE_ASSEMBLY
#if defined(SC_PROFILE)
                sc_prof_acc[38] = -1;
#endif
                assembly_shared(sm, T, 0, coop, epoch0 + step * T.nsub + sub + 1 SC_PROF_PASS);
#endif
                SC_BARL(16); SC_BARL(32); SC_BARL(8);
            }
        }
        return;
    }
    load_state(sm, T, L0, L1, a.parents, a.parent_cache, a.parent_idx, a.children, tile0, a.S);
    SC_T(0);
    for (int step = 0; step < a.nsteps; step++) {
        SC_BARL(1);
        SC_T(13);
        fk_sweep(sm, T, L0, L1, true);
        SC_T(1);
        SC_STAGE {
            SC_LANE;
            int k = tile0 + (lane & (TILE - 1));
            if (k >= a.S) k = a.S - 1;
            const float *act = a.actions + (size_t)k * (4 * T.nj);
            for (int b = 1 + slot; b < T.nb; b += LPS) pd_term_body(S, T, b, act);
        }
        SC_SYNC();
        SC_T(2);
        aba(sm, T, L0, L1, true);
        SC_T(3);
        for (int sub = 0; sub < T.nsub; sub++) {
            SC_BARL(2);
            SC_T(13);
            if (sub) { fk_sweep(sm, T, L0, L1, true); SC_T(1); }
            aba(sm, T, L0, L1, false);
            SC_T(5);
            SC_BARL(4);
            SC_T(30);
#if defined(SC_PROFILE) && defined(__CUDACC__)
            const long long sc_prof_t4 = clock64();
#endif
            // positions are those of the start of the sub-step; the ABA's scratch is free for the collision pass now
            collide(sm, T, T.self, L0, L1 SC_PROF_PASS);
            contact_cache(sm, T, L0, L1);
            SC_T(4);
            const int ncmax = tile_max_int(sm, O_NC);
#if defined(SC_PROFILE) && defined(__CUDACC__)
            sc_prof_acc[16]++; sc_prof_acc[17] += ncmax > 0; sc_prof_acc[18 + (ncmax + 1) / 2]++;
#endif
            SC_BARL(64);
            SC_T(26);
            if (ncmax > 0) contact_solve(sm, T, L0, L1, ncmax, 0 SC_PROF_PASS);      // O_TW is current: the ABA updated it
#if defined(__CUDACC__) && SC_SHARE_ASSEMBLY
#if defined(SC_PROFILE)
            sc_prof_acc[38] = clock64() - sc_prof_t4;
#endif
            assembly_shared(sm, T, ncmax, coop, epoch0 + step * T.nsub + sub + 1 SC_PROF_PASS);
            SC_T(8);
#else
            if (ncmax > 0) contact_solve(sm, T, L0, L1, ncmax, 3 SC_PROF_PASS);
#endif
            SC_BARL(16);
            SC_T(27);
#if defined(SC_PROFILE) && defined(__CUDACC__) && SC_SHARE_ASSEMBLY
            if ((threadIdx.x >> 5) == 0) {
                int *pmax = coop + 48 + 4 * ((epoch0 + step * T.nsub + sub + 1) & 1);
                for (int i = 0; i < 3; i++) sc_prof_acc[35 + i] += pmax[i];
                __syncwarp();
                if ((threadIdx.x & 31) == 0) { pmax[0] = 0; pmax[1] = 0; pmax[2] = 0; }
                sc_prof_acc[39]++;
            }
#endif
            if (ncmax > 0) { contact_solve(sm, T, L0, L1, ncmax, 1 SC_PROF_PASS); SC_T(9); }
            SC_BARL(32);
            SC_T(28);
            if (ncmax > 0) { contact_solve(sm, T, L0, L1, ncmax, 2 SC_PROF_PASS); SC_T(10); }
            SC_BARL(8);
            SC_T(29);
            integrate(sm, T, L0, L1);
            SC_T(11);
        }
    }
    features(sm, T, L0, L1);
    SC_STAGE {
        SC_LANE;
        const int s = lane & (TILE - 1), k = tile0 + s;
        if (k < a.S) {
            if (a.out_state) {
                float *dst = a.out_state + (size_t)k * SC_STATE_DIM;
                for (int i = slot; i < SC_STATE_DIM; i += LPS) dst[i] = F(state_word(T, i));
            }
            if (a.out_feat) for (int i = slot; i < 18; i += LPS) a.out_feat[(size_t)k * 18 + i] = F(O_FEAT + i);
            if (a.out_cache) {
                float *dst = a.out_cache + (size_t)k * SC_CACHE_DIM;
                for (int j = slot; j < MAXC; j += LPS) { dst[j] = (float)f2i(F(O_PCU + j)); dst[MAXC + j] = F(O_PLAM + j); }
            }
            if (a.out_overflow && slot == 0) a.out_overflow[k] = f2i(F(O_OVF));
#if defined(__CUDACC__)
            if (a.stats && slot == 0 && a.nsteps > 0) {
                atomicAdd(a.stats, (unsigned long long)(f2i(F(O_OVF)) & 0xffff));
                atomicAdd(a.stats + 1, (unsigned long long)a.nsteps * T.nsub);
                atomicAdd(a.stats + 2, (unsigned long long)(f2i(F(O_OVF)) >> 16));
            }
#endif
            if (slot == 0 && (a.out_cost || a.out_cost6)) {
                const int q = a.sample_seq ? a.sample_seq[k] : 0;
                float c[6];
                cost_terms(S, T, a.targets + (size_t)q * SC_STATE_DIM, a.tfeat + (size_t)q * 18, c);
                if (a.out_cost) a.out_cost[k] = c[0];
                if (a.out_info) for (int i = 0; i < 5; i++) a.out_info[(size_t)k * 5 + i] = c[1 + i];
                if (a.out_cost6) for (int i = 0; i < 6; i++) a.out_cost6[(size_t)k * 6 + i] = c[i];
            }
        }
    }
    SC_SYNC();
    SC_T(12);
}

}  // namespace sc

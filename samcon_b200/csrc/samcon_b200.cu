// samcon_b200.cu -- sm_100a kernels + C ABI (include/samcon_b200.h) of the SamCon sample-rollout path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py)
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <new>
#include "sc_tables.h"

using namespace sc;

// ============================================================================================================
// rollout kernel: one warp = one tile of 4 samples, 8 lanes per sample, 10 warps per CTA, the whole window in shared memory
// ============================================================================================================
constexpr int TABLE_WORDS = (sizeof(Tables) + 15) / 16 * 4;

template <int WPB>
__global__ void __launch_bounds__(32 * WPB, 1) sc_rollout_kernel(const Tables *__restrict__ gT, RolloutArgs a) {
    extern __shared__ __align__(16) float smem[];
    {
        const int4 *src = reinterpret_cast<const int4 *>(gT);
        int4 *dst = reinterpret_cast<int4 *>(smem);
        for (int i = threadIdx.x; i < TABLE_WORDS / 4; i += 32 * WPB) dst[i] = src[i];
    }
    __syncthreads();
    const Tables &T = *reinterpret_cast<const Tables *>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sm = smem + TABLE_WORDS + warp * WORDS_PER_TILE;
    const int ntiles = (a.S + TILE - 1) / TILE;
#if defined(SC_PROFILE)
    long long sc_prof_acc[SC_NPHASE] = {}, sc_prof_last = clock64();
#endif
    // Tiles are dealt evenly to the CTAs, and a CTA spreads its tiles evenly over its rounds (17 tiles: 9 + 8, not
    // 10 + 7), so the last round is not a few full CTAs next to idle SMs.  Every warp of the CTA runs every round
    // (rollout_tile has CTA-wide phase barriers); a warp without a tile passes tile0 >= S.
    const int per = ntiles / gridDim.x, extra = ntiles % gridDim.x;
    const int mine = per + ((int)blockIdx.x < extra ? 1 : 0), first = blockIdx.x * per + min((int)blockIdx.x, extra);
    const int rounds = (mine + WPB - 1) / WPB, wpr = rounds ? (mine + rounds - 1) / rounds : 0;
    for (int r = 0; r < rounds; r++) {
        const int t = r * wpr + warp;
        const bool active = warp < wpr && t < mine;
        rollout_tile(sm, T, a, active ? (first + t) * TILE : a.S, lane, lane + 1 SC_PROF_PASS);
    }
#if defined(SC_PROFILE)
    if (lane == 0 && a.prof)
        for (int i = 0; i < SC_NPHASE; i++) atomicAdd((unsigned long long *)a.prof + i, (unsigned long long)sc_prof_acc[i]);
#endif
}

// ============================================================================================================
// sample generation (SamCon.get_batch_action, /root/reference/samcon/core/samcon.py:101-142)
// ============================================================================================================
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                            uint32_t *out) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void sc_sample_gen_kernel(const float *__restrict__ target, const float *__restrict__ limits,
                                     const float *__restrict__ noise, const int *__restrict__ perm, int gaussian,
                                     uint64_t seed, uint64_t window, int S, int nj, float *__restrict__ actions) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S * nj) return;
    const int k = t / nj, j = t - k * nj;          // output row, joint
    const int src = perm ? perm[k] : k;            // row of the un-shuffled batch (np.random.shuffle, samcon.py:140)
    // target joint rotation -> intrinsic XYZ Euler angles (scipy Rotation.as_euler("XYZ"), samcon.py:123-124)
    const float *q = target + 7 + 4 * j;
    float x = q[0], y = q[1], z = q[2], w = q[3];
    const float n = rsqrtf(x * x + y * y + z * z + w * w);
    x *= n; y *= n; z *= n; w *= n;
    const float r00 = 1 - 2 * (y * y + z * z), r01 = 2 * (x * y - z * w), r02 = 2 * (x * z + y * w);
    const float r12 = 2 * (y * z - x * w), r22 = 1 - 2 * (x * x + y * y);
    float e[3];
    e[1] = atan2f(r02, sqrtf(r00 * r00 + r01 * r01));
    e[0] = atan2f(-r12, r22);
    e[2] = atan2f(-r01, r00);
    float u[3];
    if (noise) {
        for (int c = 0; c < 3; c++) u[c] = noise[(size_t)src * 3 * nj + 3 * j + c];
    } else {
        uint32_t r[4];
        philox4x32((uint32_t)t, (uint32_t)((uint64_t)t >> 32), (uint32_t)window, (uint32_t)(window >> 32), (uint32_t)seed,
                   (uint32_t)(seed >> 32), r);
        if (gaussian) {
            // Box-Muller from 4 uniforms -> 3 normals
            const float u0 = (r[0] + 1.0f) * 2.3283064e-10f, u1 = r[1] * 2.3283064e-10f;
            const float u2 = (r[2] + 1.0f) * 2.3283064e-10f, u3 = r[3] * 2.3283064e-10f;
            const float m0 = sqrtf(-2.0f * __logf(u0)), m1 = sqrtf(-2.0f * __logf(u2));
            u[0] = m0 * cospif(2.0f * u1); u[1] = m0 * sinpif(2.0f * u1); u[2] = m1 * cospif(2.0f * u3);
        } else {
            for (int c = 0; c < 3; c++) u[c] = (r[c] >> 8) * 5.9604645e-8f;     // [0,1) with 24 bits
        }
    }
    const float PI = 3.14159265358979323846f;
    for (int c = 0; c < 3; c++) {
        const float lim = limits[3 * j + c];
        // uniform: U(-v, v) (samcon.py:113-118); gaussian: N(0, v) with v the variance (samcon.py:104-111)
        e[c] += gaussian ? u[c] * sqrtf(lim) : u[c] * 2.0f * lim - lim;
        // samcon.py:130-133 wraps with `while` loops; an atan2 output plus a cfg noise range needs at most two turns, and
        // the bound keeps a non-finite input (inf limits / NaN target) from spinning the kernel
        for (int it = 0; it < 4 && e[c] > PI; it++) e[c] -= 2.0f * PI;
        for (int it = 0; it < 4 && e[c] < -PI; it++) e[c] += 2.0f * PI;
    }
    // intrinsic XYZ Euler -> quaternion = qx (x) qy (x) qz (scipy from_euler("XYZ").as_quat(), samcon.py:136-137)
    float sa, ca, sb, cb, sc_, cc;
    sincosf(0.5f * e[0], &sa, &ca); sincosf(0.5f * e[1], &sb, &cb); sincosf(0.5f * e[2], &sc_, &cc);
    float *o = actions + (size_t)k * 4 * nj + 4 * j;
    o[0] = sa * cb * cc + ca * sb * sc_;
    o[1] = ca * sb * cc - sa * cb * sc_;
    o[2] = ca * cb * sc_ + sa * sb * cc;
    o[3] = ca * cb * cc - sa * sb * sc_;
}

// ============================================================================================================
// elite selection: segmented radix select of the E smallest (cost, index) keys + bitonic sort of the winners
// (Trajectory.select('greedy'), /root/reference/samcon/core/trajectory.py:108-112)
// ============================================================================================================
__device__ __forceinline__ unsigned long long sel_key(float c, int i) {
    uint32_t b = __float_as_uint(c);
    if (c != c) b = 0xFFFFFFFFu;                    // NaN sorts last
    else b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)i;
}

constexpr int SEL_THREADS = 1024;

__global__ void __launch_bounds__(SEL_THREADS) sc_select_kernel(const float *__restrict__ cost, const int *__restrict__ seg,
                                                                int E, int Epad, unsigned long long *__restrict__ scratch,
                                                                int *__restrict__ elite_idx, float *__restrict__ elite_cost) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_rank, s_count;
    const int sg = blockIdx.x, base = seg[sg], n = seg[sg + 1] - base, tid = threadIdx.x;
    const float *c = cost + base;
    unsigned long long *buf = scratch + (size_t)sg * Epad;
    const int Ee = E < n ? E : n;   // a segment shorter than E yields its n entries, then -1 / +inf padding
    if (tid == 0) { s_prefix = 0; s_rank = Ee > 0 ? Ee - 1 : 0; s_count = 0; }
    __syncthreads();
    // MSD radix select over the 64-bit keys: after 8 digit passes s_prefix is the E-th smallest key
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix, mask = pass ? (~0ull << (shift + 8)) : 0ull;
        for (int i = tid; i < n; i += SEL_THREADS) {
            const unsigned long long k = sel_key(c[i], i);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int r = s_rank, acc = 0;
            int d = 0;
            for (; d < 256; d++) { if (acc + hist[d] > r) break; acc += hist[d]; }
            s_rank = r - acc;
            s_prefix = prefix | ((unsigned long long)d << shift);
        }
        __syncthreads();
    }
    const unsigned long long thr = s_prefix;
    for (int i = tid; i < Epad; i += SEL_THREADS) buf[i] = ~0ull;
    __syncthreads();
    for (int i = tid; i < n; i += SEL_THREADS) {
        const unsigned long long k = sel_key(c[i], i);
        if (Ee > 0 && k <= thr) buf[atomicAdd(&s_count, 1u)] = k;
    }
    __syncthreads();
    // bitonic sort of Epad keys (ascending) in global scratch; one CTA, so __syncthreads orders the passes
    for (int k2 = 2; k2 <= Epad; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Epad; i += SEL_THREADS) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long a = buf[i], b = buf[l];
                    const bool up = (i & k2) == 0;
                    if ((a > b) == up) { buf[i] = b; buf[l] = a; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < E; i += SEL_THREADS) {
        const int li = (int)(uint32_t)buf[i];
        elite_idx[(size_t)sg * E + i] = i < Ee ? base + li : -1;
        if (elite_cost) elite_cost[(size_t)sg * E + i] = i < Ee ? c[li] : __int_as_float(0x7f800000);
    }
}

__global__ void sc_gather_rows_kernel(const float *__restrict__ src, const int *__restrict__ idx, int n, int width,
                                      float *__restrict__ dst) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * width) return;
    const int r = (int)(t / width), cidx = (int)(t - (size_t)r * width);
    dst[t] = src[(size_t)idx[r] * width + cidx];
}

__global__ void sc_iota_kernel(int *p, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = t;
}

// ============================================================================================================
// C ABI
// ============================================================================================================
struct sc_context {
    int device;
    Tables h_tables;
    Tables *d_tables;
    float *d_tfeat; size_t tfeat_cap;
    unsigned long long *d_sel; size_t sel_cap;
    int *d_iota; size_t iota_cap;
    int smem_bytes, grid_cap;
    long long launches;
    long long *d_prof;
    char err[512];
};

static int sc_fail(sc_context *h, int code, const char *fmt, ...) {
    if (h) { va_list ap; va_start(ap, fmt); vsnprintf(h->err, sizeof h->err, fmt, ap); va_end(ap); }
    return code;
}
#define SC_CUDA(h, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return sc_fail(h, SC_ECUDA, "%s: %s", #call, cudaGetErrorString(e_)); } while (0)

#ifndef SC_WPB
#define SC_WPB 10
#endif
constexpr int WPB = SC_WPB;   // warps (tiles of 4 samples) per CTA: 40 samples x 5.4 KB + tables fill the 227 KB of one SM

extern "C" int sc_create(sc_handle *out, const sc_model *model, int device) {
    if (!out) return SC_EINVAL;
    *out = nullptr;
    sc_context *h = new (std::nothrow) sc_context();
    if (!h) return SC_ENOMEM;
    memset(h, 0, sizeof *h);
    h->device = device;
    *out = h;   // returned even on failure so that sc_last_error() can explain; caller still sc_destroy()s it
    int rc = build_tables(model, &h->h_tables, h->err, sizeof h->err);
    if (rc) return rc;
    SC_CUDA(h, cudaSetDevice(device));
    SC_CUDA(h, cudaMalloc(&h->d_tables, TABLE_WORDS * 4));
    SC_CUDA(h, cudaMemset(h->d_tables, 0, TABLE_WORDS * 4));
    SC_CUDA(h, cudaMemcpy(h->d_tables, &h->h_tables, sizeof(Tables), cudaMemcpyHostToDevice));
    h->smem_bytes = (TABLE_WORDS + WPB * WORDS_PER_TILE) * 4;
    SC_CUDA(h, cudaFuncSetAttribute(sc_rollout_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    int nsm = 0, per_sm = 0;
    SC_CUDA(h, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    SC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sc_rollout_kernel<WPB>, 32 * WPB, h->smem_bytes));
    if (per_sm < 1) return sc_fail(h, SC_ECUDA, "rollout kernel does not fit an SM (%d B shared)", h->smem_bytes);
    h->grid_cap = nsm * per_sm;
    return SC_OK;
}

extern "C" int sc_destroy(sc_handle h) {
    if (!h) return SC_EINVAL;
    if (h->d_tables) { cudaSetDevice(h->device); cudaFree(h->d_tables); cudaFree(h->d_tfeat); cudaFree(h->d_sel); cudaFree(h->d_iota); }
    delete h;
    return SC_OK;
}

extern "C" const char *sc_last_error(sc_handle h) { return h ? h->err : "null handle"; }
extern "C" int64_t sc_launch_count(sc_handle h) { return h ? h->launches : 0; }
extern "C" int sc_rollout_config(sc_handle h, int *smem_bytes, int *samples_per_cta, int *threads_per_cta) {
    if (!h) return SC_EINVAL;
    if (smem_bytes) *smem_bytes = h->smem_bytes;
    if (samples_per_cta) *samples_per_cta = WPB * TILE;
    if (threads_per_cta) *threads_per_cta = WPB * 32;
    return SC_OK;
}

#if defined(SC_PROFILE)
// developer build only (tools/phase_profile.py): read and reset the per-phase cycle counters
extern "C" int sc_prof_read(sc_handle h, long long *out) {
    if (!h || !h->d_prof) return SC_EINVAL;
    cudaDeviceSynchronize();
    cudaMemcpy(out, h->d_prof, SC_NPHASE * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemset(h->d_prof, 0, SC_NPHASE * sizeof(long long));
    return SC_OK;
}
#endif

template <class T>
static int sc_reserve(sc_context *h, T **p, size_t *cap, size_t need) {
    if (*cap >= need) return SC_OK;
    // growing a scratch buffer: the old one may still be in use by queued work
    SC_CUDA(h, cudaDeviceSynchronize());
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    SC_CUDA(h, cudaMalloc(p, need * sizeof(T)));
    *cap = need;
    return SC_OK;
}

static int sc_launch_rollout(sc_context *h, const RolloutArgs &a, cudaStream_t st) {
    const int ntiles = (a.S + TILE - 1) / TILE;
    // all SMs even when there are few tiles (walk.yml's own nSample = 2000 is 500 tiles: 3-4 warps on each of 148 SMs
    // beat 10 warps on 50 of them); the kernel deals the tiles evenly
    const int grid = ntiles < h->grid_cap ? ntiles : h->grid_cap;
#if defined(SC_PROFILE)
    RolloutArgs ap = a;
    if (!h->d_prof) { cudaMalloc(&h->d_prof, SC_NPHASE * sizeof(long long)); cudaMemset(h->d_prof, 0, SC_NPHASE * sizeof(long long)); }
    ap.prof = a.nsteps > 0 ? h->d_prof : nullptr;
    sc_rollout_kernel<WPB><<<grid, 32 * WPB, h->smem_bytes, st>>>(h->d_tables, ap);
#else
    sc_rollout_kernel<WPB><<<grid, 32 * WPB, h->smem_bytes, st>>>(h->d_tables, a);
#endif
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

// features of `n` states (rows of `states`) -> d_tfeat[n,18]
static int sc_features(sc_context *h, const float *states, int n, cudaStream_t st) {
    int rc = sc_reserve(h, &h->d_tfeat, &h->tfeat_cap, (size_t)n * 18);
    if (rc) return rc;
    RolloutArgs f = {};
    f.parents = states; f.children = 1; f.S = n; f.nsteps = 0; f.out_feat = h->d_tfeat;
    return sc_launch_rollout(h, f, st);
}

extern "C" int sc_window(sc_handle h, const float *parents, int E, const int32_t *parent_idx, int children,
                         const float *actions, const float *targets, int nseq, const int32_t *sample_seq, int S, int nsteps,
                         float *out_state, float *out_cost, float *out_info, void *stream) {
    if (!h) return SC_EINVAL;
    if (!parents || !actions || !targets || S < 1 || E < 1 || nseq < 1 || nsteps < 0)
        return sc_fail(h, SC_EINVAL, "sc_window: bad arguments (S=%d E=%d nseq=%d nsteps=%d)", S, E, nseq, nsteps);
    if (!parent_idx && (children < 1 || (long long)E * children < S))
        return sc_fail(h, SC_EINVAL, "sc_window: %d parents x %d children cannot seed %d samples", E, children, S);
    if (nseq > 1 && !sample_seq) return sc_fail(h, SC_EINVAL, "sc_window: nseq=%d needs sample_seq", nseq);
    cudaStream_t st = (cudaStream_t)stream;
    SC_CUDA(h, cudaSetDevice(h->device));
    int rc = sc_features(h, targets, nseq, st);
    if (rc) return rc;
    RolloutArgs a = {};
    a.parents = parents; a.parent_idx = parent_idx; a.children = children > 0 ? children : 1; a.actions = actions;
    a.targets = targets; a.tfeat = h->d_tfeat; a.sample_seq = sample_seq; a.S = S; a.nsteps = nsteps;
    a.out_state = out_state; a.out_cost = out_cost; a.out_info = out_info;
    return sc_launch_rollout(h, a, st);
}

extern "C" int sc_step(sc_handle h, float *state, const float *actions, int S, int nsteps, void *stream) {
    if (!h) return SC_EINVAL;
    if (!state || !actions || S < 1 || nsteps < 0) return sc_fail(h, SC_EINVAL, "sc_step: bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    RolloutArgs a = {};
    a.parents = state; a.children = 1; a.actions = actions; a.S = S; a.nsteps = nsteps; a.out_state = state;
    return sc_launch_rollout(h, a, (cudaStream_t)stream);
}

extern "C" int sc_cost(sc_handle h, const float *sim, const float *kin, int S, float *out, void *stream) {
    if (!h) return SC_EINVAL;
    if (!sim || !kin || !out || S < 1) return sc_fail(h, SC_EINVAL, "sc_cost: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SC_CUDA(h, cudaSetDevice(h->device));
    int rc = sc_features(h, kin, S, st);
    if (rc) return rc;
    rc = sc_reserve(h, &h->d_iota, &h->iota_cap, (size_t)S);
    if (rc) return rc;
    sc_iota_kernel<<<(S + 255) / 256, 256, 0, st>>>(h->d_iota, S);
    h->launches++;
    // cost lands in out[:,0] via out_cost stride trick: run with out_cost/out_info on a packed [S,6] buffer
    RolloutArgs a = {};
    a.parents = sim; a.children = 1; a.actions = sim; a.targets = kin; a.tfeat = h->d_tfeat; a.sample_seq = h->d_iota;
    a.S = S; a.nsteps = 0; a.out_cost = nullptr; a.out_info = nullptr; a.out_feat = nullptr;
    a.out_cost6 = out;
    return sc_launch_rollout(h, a, st);
}

extern "C" int sc_sample_gen(sc_handle h, const float *target, const float *limits, const float *noise, const int32_t *perm,
                             int gaussian, uint64_t seed, uint64_t window, int S, float *actions, void *stream) {
    if (!h) return SC_EINVAL;
    if (!target || !limits || !actions || S < 1) return sc_fail(h, SC_EINVAL, "sc_sample_gen: bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    const int nj = h->h_tables.nj, n = S * nj;
    sc_sample_gen_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(target, limits, noise, perm, gaussian, seed, window, S,
                                                                             nj, actions);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_select(sc_handle h, const float *cost, const int32_t *seg_offsets, int nseg, int E, int32_t *elite_idx,
                         float *elite_cost, void *stream) {
    if (!h) return SC_EINVAL;
    if (!cost || !seg_offsets || !elite_idx || nseg < 1 || E < 1) return sc_fail(h, SC_EINVAL, "sc_select: bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    int Epad = 2;
    while (Epad < E) Epad <<= 1;
    int rc = sc_reserve(h, &h->d_sel, &h->sel_cap, (size_t)nseg * Epad);
    if (rc) return rc;
    sc_select_kernel<<<nseg, SEL_THREADS, 0, (cudaStream_t)stream>>>(cost, seg_offsets, E, Epad, h->d_sel, elite_idx, elite_cost);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_gather_rows(sc_handle h, const float *src, const int32_t *idx, int n, int width, float *dst, void *stream) {
    if (!h) return SC_EINVAL;
    if (!src || !idx || !dst || n < 1 || width < 1) return sc_fail(h, SC_EINVAL, "sc_gather_rows: bad arguments");
    SC_CUDA(h, cudaSetDevice(h->device));
    const size_t tot = (size_t)n * width;
    sc_gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, width, dst);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

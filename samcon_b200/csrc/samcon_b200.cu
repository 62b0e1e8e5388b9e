// samcon_b200.cu -- sm_100a kernels + C ABI (include/samcon_b200.h) of the SamCon sample-rollout path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py)
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <new>
#include "sc_tables.h"

using namespace sc;
namespace cg = cooperative_groups;

// ============================================================================================================
// rollout kernel: one warp = one tile of 4 samples, 8 lanes per sample, 10 warps per CTA, the whole window in shared memory
// ============================================================================================================
constexpr int TABLE_WORDS = (sizeof(Tables) + 15) / 16 * 4;

// WPB warps with a tile each (what shared memory holds) + up to HELPER_WARPS warps without one, which only take column calls
// of the shared contact assembly (the register file has room for them: 12 warps x 32 x 168 registers)
constexpr int HELPER_WARPS = 2;
template <int WPB, class TT, int BM>
__global__ void __launch_bounds__(32 * (WPB + HELPER_WARPS), 1) sc_rollout_kernel(const Tables *__restrict__ gT, RolloutArgs a) {
    extern __shared__ __align__(16) float smem[];
    {
        const int4 *src = reinterpret_cast<const int4 *>(gT);
        int4 *dst = reinterpret_cast<int4 *>(smem);
        for (int i = threadIdx.x; i < TABLE_WORDS / 4; i += blockDim.x) dst[i] = src[i];
    }
    // behind the tiles: the bookkeeping of the CTA's shared contact assembly (sc_core.cuh: assembly_shared), epochs start at 1
    int *coop = reinterpret_cast<int *>(smem + TABLE_WORDS + a.tile_warps * WORDS_PER_TILE);
    for (int i = threadIdx.x; i < COOP_WORDS; i += blockDim.x) coop[i] = 0;
    __syncthreads();
    const TT &T = *reinterpret_cast<const TT *>(smem);
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
    // WPB is the build's ceiling; the launch brings only as many warps as a round needs (blockDim.x / 32 = wpr of the fullest
    // CTA, see sc_launch_rollout), so no warp sits at the phase barriers without a tile round after round
    const int nw = a.tile_warps;                    // (the warps behind them are helpers: never active, no tile memory)
    const int rounds = (mine + nw - 1) / nw, wpr = rounds ? (mine + rounds - 1) / rounds : 0;
    for (int r = 0; r < rounds; r++) {
        const int t = r * wpr + warp;
        const bool active = warp < wpr && t < mine;
        rollout_tile<TT, BM>(sm, T, a, active ? (first + t) * TILE : a.S, lane, lane + 1, coop, r * a.nsteps * T.nsub SC_PROF_PASS);
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

// Sequence q of nseq owns rows [q S, (q+1) S): its target is targets[q], its Philox stream is keyed by seed + q and counted
// from its first row, so a batch of sequences draws what each sequence would draw alone.
__global__ void sc_sample_gen_kernel(const float *__restrict__ targets, const float *__restrict__ limits,
                                     const float *__restrict__ noise, const int *__restrict__ perm, int gaussian,
                                     uint64_t seed0, uint64_t window, int S, int nseq, int nj, float *__restrict__ actions) {
    const long long tg = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tg >= (long long)nseq * S * nj) return;
    const int seq = (int)(tg / ((long long)S * nj));
    const int t = (int)(tg - (long long)seq * S * nj);
    const uint64_t seed = seed0 + (uint64_t)seq;
    const int k = t / nj, j = t - k * nj;          // output row in the sequence, joint
    const size_t row0 = (size_t)seq * S;
    const int src = perm ? perm[row0 + k] : k;     // row of the un-shuffled batch (np.random.shuffle, samcon.py:140)
    // target joint rotation -> intrinsic XYZ Euler angles (scipy Rotation.as_euler("XYZ"), samcon.py:123-124)
    const float *q = targets + (size_t)seq * SC_STATE_DIM + 7 + 4 * j;
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
        for (int c = 0; c < 3; c++) u[c] = noise[(row0 + src) * 3 * nj + 3 * j + c];
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
    float *o = actions + (row0 + k) * 4 * nj + 4 * j;
    o[0] = sa * cb * cc + ca * sb * sc_;
    o[1] = ca * sb * cc - sa * cb * sc_;
    o[2] = ca * cb * sc_ + sa * sb * cc;
    o[3] = ca * cb * cc - sa * sb * sc_;
}

// ============================================================================================================
// elite selection: segmented radix select of the E smallest (cost, index) keys + bitonic sort of the winners
// (Trajectory.select('greedy'), /root/reference/samcon/core/trajectory.py:108-112)
//
// One cooperative launch for all segments.  Every segment is cut into chunks of `chunk` costs, one CTA per chunk (a
// 10^6-sample window spreads over the whole GPU, 64 sequences of 2000 get a CTA each).  Eight MSD passes over the 64-bit
// key (order-preserving cost bits, index in the segment): each CTA histograms the digit of its chunk's keys that still
// match the prefix in shared memory, adds the 256 bins to the segment's global histogram, and after a grid-wide
// barrier every CTA of the segment scans those bins with one warp (identical result everywhere, so nobody has to
// publish it).  The E-th smallest key is then a threshold: the chunks append their keys <= it to the segment's list
// (order does not matter, the keys are unique), and the list is sorted -- in shared memory by the segment's first CTA
// when it fits, else by a grid-wide bitonic network whose wide strides go through global memory.
// ============================================================================================================
__device__ __forceinline__ unsigned long long sel_key(float c, int i) {
    uint32_t b = __float_as_uint(c);
    if (c != c) b = 0xFFFFFFFFu;                    // NaN sorts last
    else b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)i;
}

constexpr int SEL_THREADS = 512;
constexpr int SEL_TILE = 4096;                      // keys sorted in one CTA's shared memory (32 KB)

struct SelectArgs {
    const float *cost;
    const int *seg;
    int nseg, E, Epad, chunk;
    unsigned long long *buf;                        // [nseg, Epad] keys of the winners
    unsigned int *ghist;                            // [8, nseg, 256] zeroed by the caller
    unsigned int *gcount;                           // [nseg] zeroed by the caller
    int *elite_idx;
    float *elite_cost;
};

__device__ __forceinline__ void sel_cmpx(unsigned long long &a, unsigned long long &b, bool up) {
    if ((a > b) == up) { const unsigned long long t = a; a = b; b = t; }
}

// bitonic stages j = j0, j0/2, .. 1 of merge size k on a tile held in shared memory; gbase = index of tile[0] in its block
__device__ __forceinline__ void sel_local_stages(unsigned long long *tile, int n, int gbase, int k, int j0) {
    for (int j = j0; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < n / 2; t += SEL_THREADS) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
            sel_cmpx(tile[i], tile[l], ((gbase + i) & k) == 0);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SEL_THREADS) sc_select_kernel(SelectArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned long long tile[];
    __shared__ unsigned int hist[256];
    __shared__ int s_seg, s_begin, s_end, s_first;
    __shared__ unsigned long long s_prefix;
    __shared__ unsigned int s_rank;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int b = blockIdx.x, sg = -1, begin = 0, end = 0;
        for (int g = 0; g < a.nseg; g++) {
            const int n = a.seg[g + 1] - a.seg[g], nch = n > 0 ? (n + a.chunk - 1) / a.chunk : 1;
            if (b < nch) { sg = g; begin = a.seg[g] + b * a.chunk; end = min(a.seg[g + 1], begin + a.chunk); break; }
            b -= nch;
        }
        s_seg = sg; s_begin = begin; s_end = end; s_first = b == 0;
    }
    __syncthreads();
    const int sg = s_seg, begin = s_begin, end = s_end;
    const int base = sg >= 0 ? a.seg[sg] : 0, n = sg >= 0 ? a.seg[sg + 1] - base : 0;
    const int Ee = a.E < n ? a.E : n;             // a segment shorter than E yields its n entries, then -1 / +inf padding
    unsigned long long prefix = 0;
    unsigned int rank = Ee > 0 ? Ee - 1 : 0;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        const unsigned long long mask = pass ? (~0ull << (shift + 8)) : 0ull;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int i = begin + tid; i < end; i += SEL_THREADS) {
            const unsigned long long k = sel_key(a.cost[i], i - base);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255], 1u);
        }
        __syncthreads();
        if (sg >= 0 && tid < 256 && hist[tid]) atomicAdd(&a.ghist[((size_t)pass * a.nseg + sg) * 256 + tid], hist[tid]);
        grid.sync();
        if (sg >= 0 && tid < 32) {
            // bins 8 lane .. 8 lane + 7 per lane; exclusive scan of the lane sums finds the lane, then the bin
            const unsigned int *gh = a.ghist + ((size_t)pass * a.nseg + sg) * 256 + 8 * tid;
            unsigned int v[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { v[q] = __ldcg(gh + q); sum += v[q]; }
            unsigned int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += u; }
            const unsigned int exc = inc - sum;
            const bool mine = exc <= rank && rank < inc;
            const unsigned int who = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                unsigned int acc = exc;
                int d = 0;
                for (; d < 8; d++) { if (acc + v[d] > rank) break; acc += v[d]; }
                s_prefix = prefix | ((unsigned long long)(8 * tid + d) << shift);
                s_rank = rank - acc;
            } else if (who == 0 && tid == 0) { s_prefix = prefix | (255ull << shift); s_rank = 0; }   // empty segment
        }
        __syncthreads();
        prefix = s_prefix; rank = s_rank;
        __syncthreads();
    }
    // ---- the Ee keys <= prefix are the winners
    if (sg >= 0 && Ee > 0)
        for (int i = begin + tid; i < end; i += SEL_THREADS) {
            const unsigned long long k = sel_key(a.cost[i], i - base);
            if (k <= prefix) a.buf[(size_t)sg * a.Epad + atomicAdd(&a.gcount[sg], 1u)] = k;
        }
    if (sg >= 0 && s_first)
        for (int i = Ee + tid; i < a.Epad; i += SEL_THREADS) a.buf[(size_t)sg * a.Epad + i] = ~0ull;
    grid.sync();
    // ---- sort
    if (a.Epad <= SEL_TILE) {
        if (sg >= 0 && s_first) {
            unsigned long long *src = a.buf + (size_t)sg * a.Epad;
            for (int i = tid; i < a.Epad; i += SEL_THREADS) tile[i] = __ldcg(src + i);
            __syncthreads();
            for (int k = 2; k <= a.Epad; k <<= 1) sel_local_stages(tile, a.Epad, 0, k, k >> 1);
            for (int i = tid; i < a.E; i += SEL_THREADS) {
                const int li = (int)(uint32_t)tile[i];
                a.elite_idx[(size_t)sg * a.E + i] = i < Ee ? base + li : -1;
                if (a.elite_cost) a.elite_cost[(size_t)sg * a.E + i] = i < Ee ? a.cost[base + li] : __int_as_float(0x7f800000);
            }
        }
        return;
    }
    // wide lists: every CTA works on tiles of the whole [nseg * Epad] array; a block of Epad keys is one bitonic problem
    const size_t total = (size_t)a.nseg * a.Epad;
    const int ntile = (int)(total / SEL_TILE);
    for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
        unsigned long long *p = a.buf + (size_t)t * SEL_TILE;
        const int gbase = (int)(((size_t)t * SEL_TILE) & (a.Epad - 1));
        for (int i = tid; i < SEL_TILE; i += SEL_THREADS) tile[i] = __ldcg(p + i);
        __syncthreads();
        for (int k = 2; k <= SEL_TILE; k <<= 1) sel_local_stages(tile, SEL_TILE, gbase, k, k >> 1);
        for (int i = tid; i < SEL_TILE; i += SEL_THREADS) p[i] = tile[i];
        __syncthreads();
    }
    grid.sync();
    for (int k = 2 * SEL_TILE; k <= a.Epad; k <<= 1) {
        for (int j = k >> 1; j >= SEL_TILE; j >>= 1) {
            for (size_t t = (size_t)blockIdx.x * SEL_THREADS + tid; t < total / 2; t += (size_t)gridDim.x * SEL_THREADS) {
                const size_t i = ((t & ~(size_t)(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                unsigned long long x = __ldcg(a.buf + i), y = __ldcg(a.buf + l);
                const bool up = ((i & (a.Epad - 1)) & k) == 0;
                if ((x > y) == up) { a.buf[i] = y; a.buf[l] = x; }
            }
            grid.sync();
        }
        for (int t = blockIdx.x; t < ntile; t += gridDim.x) {
            unsigned long long *p = a.buf + (size_t)t * SEL_TILE;
            const int gbase = (int)(((size_t)t * SEL_TILE) & (a.Epad - 1));
            for (int i = tid; i < SEL_TILE; i += SEL_THREADS) tile[i] = __ldcg(p + i);
            __syncthreads();
            sel_local_stages(tile, SEL_TILE, gbase, k, SEL_TILE >> 1);
            for (int i = tid; i < SEL_TILE; i += SEL_THREADS) p[i] = tile[i];
            __syncthreads();
        }
        grid.sync();
    }
    for (size_t t = (size_t)blockIdx.x * SEL_THREADS + tid; t < (size_t)a.nseg * a.E; t += (size_t)gridDim.x * SEL_THREADS) {
        const int g = (int)(t / a.E), i = (int)(t - (size_t)g * a.E);
        const int gb = a.seg[g], gn = a.seg[g + 1] - gb, ge = a.E < gn ? a.E : gn;
        const int li = (int)(uint32_t)__ldcg(a.buf + (size_t)g * a.Epad + i);
        a.elite_idx[t] = i < ge ? gb + li : -1;
        if (a.elite_cost) a.elite_cost[t] = i < ge ? a.cost[gb + li] : __int_as_float(0x7f800000);
    }
}

__global__ void sc_gather_rows_kernel(const float *__restrict__ src, const int *__restrict__ idx, int n, int width,
                                      float *__restrict__ dst) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * width) return;
    const int r = (int)(t / width), cidx = (int)(t - (size_t)r * width);
    const int i = idx[r];
    dst[t] = i >= 0 ? src[(size_t)i * width + cidx] : 0.0f;      // -1 = sc_select's padding of a short segment: a zero row
}

// rows fetched out of other GPUs' buffers: peer_base[p] is the device address of peer p's [*, width] array (mapped into this
// process by CUDA VMM / symmetric memory); the loads travel over NVLink / NVSwitch, coalesced along the row
__global__ void sc_gather_rows_peer_kernel(const unsigned long long *__restrict__ peer_base, const int *__restrict__ src_peer,
                                           const int *__restrict__ src_row, int n, int width, float *__restrict__ dst) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * width) return;
    const int r = (int)(t / width), cidx = (int)(t - (size_t)r * width);
    const float *src = reinterpret_cast<const float *>(peer_base[src_peer[r]]);
    dst[t] = __ldcv(src + (size_t)src_row[r] * width + cidx);      // volatile load: never a stale line of an earlier window
}

__global__ void sc_iota_kernel(int *p, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = t;
}

// ============================================================================================================
// reference-motion targets (AmassMotionData.getSpecTimeState, /root/reference/data_loaders/amass_motion_data.py:48-109,
// with compute_lin_vel / compute_ang_vel / compute_ang_vel_rel of /root/reference/utils/bullet_utils.py:9-45):
// the 132-float target state of every (sequence, query time) pair in one launch.  One thread per (sequence, time,
// rotation): thread r = 0 handles the root (position, linear velocity, world-frame angular velocity), r = 1..nj a joint
// (local-frame angular velocity).  Frame lookup runs in double like the reference's Python floats -- a time that lands
// on a frame boundary must pick the frame the reference picks -- the quaternion arithmetic in float.
// ============================================================================================================
__device__ __forceinline__ void q_mul(const float *a, const float *b, float *o) {
    o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
    o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
    o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
}

__global__ void sc_motion_states_kernel(const float *__restrict__ poses, const int *__restrict__ frame_off,
                                        const double *__restrict__ frame_dt, const double *__restrict__ times, int nseq, int nt,
                                        int nj, float *__restrict__ out) {
    const int nr = nj + 1;
    const long long tg = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tg >= (long long)nseq * nt * nr) return;
    const int r = (int)(tg % nr), it = (int)((tg / nr) % nt), q = (int)(tg / ((long long)nr * nt));
    const int f0 = frame_off[q], T = frame_off[q + 1] - f0, pd = 3 + 4 * nr;
    const double dur = frame_dt[q], cycle = dur * (T - 1), t = times[it];
    const double count = floor(t / cycle);
    double ft = t - count * cycle;
    if (ft < 0) ft += cycle;
    const int frame = (int)(ft / dur);
    const int next = frame + 1 >= T ? frame : frame + 1;
    const float frac = (float)((ft - frame * dur) / dur), idt = (float)(1.0 / dur);
    const float *A = poses + (size_t)(f0 + frame) * pd, *B = poses + (size_t)(f0 + next) * pd;
    float *o = out + ((size_t)q * nt + it) * SC_STATE_DIM;
    const float *qa = A + 3 + 4 * r, *qb = B + 3 + 4 * r;
    // getQuaternionSlerp: shortest arc, weights sin((1-f) theta) / sin(theta) and +-sin(f theta) / sin(theta) with theta the
    // angle between the two 4-vectors; (anti)parallel inputs return the start orientation.  The reference takes theta from
    // acos(|<q0,q1>|) in double; in float that loses everything below 7e-4 rad (consecutive frames!), so theta comes from the
    // chord |+-q1 - q0| = 2 sin(theta / 2) instead, and the weights go over into their limit 1 - f, +-f for tiny angles.
    const float na = rsqrtf(qa[0] * qa[0] + qa[1] * qa[1] + qa[2] * qa[2] + qa[3] * qa[3]);
    const float nb = rsqrtf(qb[0] * qb[0] + qb[1] * qb[1] + qb[2] * qb[2] + qb[3] * qb[3]);
    const float prod = (qa[0] * qb[0] + qa[1] * qb[1] + qa[2] * qb[2] + qa[3] * qb[3]) * na * nb;
    const float sign = prod < 0 ? -1.0f : 1.0f;
    float ch2 = 0.0f;
    for (int c = 0; c < 4; c++) { const float d = sign * qb[c] * nb - qa[c] * na; ch2 += d * d; }
    const float theta = 2.0f * asinf(fminf(1.0f, 0.5f * sqrtf(ch2)));
    float s0 = 1.0f - frac, s1 = sign * frac;
    if (theta > 1e-3f) {
        const float d = 1.0f / sinf(theta);
        s0 = sinf((1.0f - frac) * theta) * d;
        s1 = sign * sinf(frac * theta) * d;
    }
    if (ch2 == 0.0f) { s0 = 1.0f; s1 = 0.0f; }
    float *oq = o + 3 + 4 * r;
    for (int c = 0; c < 4; c++) oq[c] = qa[c] * s0 + qb[c] * s1;
    // angular velocity: axis-angle of q1 (x) conj(q0) (root, world frame) or conj(q0) (x) q1 (joints, local frame), over the
    // key-frame duration; getAxisAngleFromQuaternion has no shortest-path flip
    const float qc[4] = {-qa[0], -qa[1], -qa[2], qa[3]};
    float dq[4];
    if (r == 0) q_mul(qb, qc, dq); else q_mul(qc, qb, dq);
    // (angle 2 acos(w), axis xyz / sqrt(1 - w^2) in the reference's doubles; in float the same numbers come out accurately as
    // 2 atan2(|xyz|, w) and xyz / |xyz| -- acosf loses half its digits next to w = 1, where consecutive frames live)
    const float s2 = dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2], sn = sqrtf(s2), ang = 2.0f * atan2f(sn, dq[3]);
    float ax[3] = {1.0f, 0.0f, 0.0f};
    if (s2 >= 1e-14f) { const float k = 1.0f / sn; ax[0] = dq[0] * k; ax[1] = dq[1] * k; ax[2] = dq[2] * k; }
    float *ow = o + 3 + 4 * nr + 3 + 3 * r;            // root omega sits after the root's linear velocity, joints follow
    for (int c = 0; c < 3; c++) ow[c] = ax[c] * ang * idt;
    if (r == 0) {
        const double off_scale = count;
        const float *P0 = poses + (size_t)f0 * pd, *PL = poses + (size_t)(f0 + T - 1) * pd;
        for (int c = 0; c < 3; c++) {
            o[c] = A[c] + frac * (B[c] - A[c]) + (float)(off_scale * ((double)PL[c] - (double)P0[c]));
            o[3 + 4 * nr + c] = (B[c] - A[c]) * idt;
        }
    }
}

// ============================================================================================================
// measured FP32 peak: the denominator of the roofline this path is held against (MEASURED_PEAKS.json has HBM and bf16
// tensor figures only).  Eight independent FFMA chains per thread, 8 warps per scheduler: nothing but the FMA pipe.
// ============================================================================================================
__global__ void __launch_bounds__(256) sc_ffma_peak_kernel(float *out, int iters, float b, float c) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; k++) a[k] = 1e-3f * (threadIdx.x + k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) a[k] = fmaf(a[k], b, c);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ============================================================================================================
// C ABI
// ============================================================================================================
// every entry point runs on the handle's device and puts the caller's current device back before returning
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess; else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define SC_ON_DEVICE(h) DeviceGuard guard_((h)->device); if (!guard_.ok) return sc_fail(h, SC_ECUDA, "cudaSetDevice(%d) failed", (h)->device)

struct sc_context {
    int device;
    Tables h_tables;
    Tables *d_tables;
    float *d_tfeat; size_t tfeat_cap;
    unsigned long long *d_sel; size_t sel_cap;      // select scratch: winners' keys, then histograms and counters
    int sel_grid_cap;                               // co-resident CTAs of sc_select_kernel (cooperative launch)
    int *d_iota; size_t iota_cap;
    int smem_bytes, grid_cap;
    long long launches;
    long long *d_prof;
    unsigned long long *d_stats;    // [3] contact statistics, see sc_contact_stats
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
    SC_ON_DEVICE(h);
    SC_CUDA(h, cudaMalloc(&h->d_stats, 3 * sizeof(unsigned long long)));
    SC_CUDA(h, cudaMemset(h->d_stats, 0, 3 * sizeof(unsigned long long)));
    SC_CUDA(h, cudaMalloc(&h->d_tables, TABLE_WORDS * 4));
    SC_CUDA(h, cudaMemset(h->d_tables, 0, TABLE_WORDS * 4));
    SC_CUDA(h, cudaMemcpy(h->d_tables, &h->h_tables, sizeof(Tables), cudaMemcpyHostToDevice));
    h->smem_bytes = (TABLE_WORDS + WPB * WORDS_PER_TILE + COOP_WORDS) * 4;
    SC_CUDA(h, cudaFuncSetAttribute(sc_rollout_kernel<WPB, TablesCapsule, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    SC_CUDA(h, cudaFuncSetAttribute(sc_rollout_kernel<WPB, TablesBox, 31>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    SC_CUDA(h, cudaFuncSetAttribute(sc_rollout_kernel<WPB, TablesCapsule, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    SC_CUDA(h, cudaFuncSetAttribute(sc_rollout_kernel<WPB, TablesBox, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
    int nsm = 0, per_sm = 0;
    SC_CUDA(h, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    SC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sc_rollout_kernel<WPB, TablesBox, 31>, 32 * WPB, h->smem_bytes));
    if (per_sm < 1) return sc_fail(h, SC_ECUDA, "rollout kernel does not fit an SM (%d B shared)", h->smem_bytes);
    h->grid_cap = nsm * per_sm;
    int sel_per_sm = 0;
    SC_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sel_per_sm, sc_select_kernel, SEL_THREADS, SEL_TILE * sizeof(unsigned long long)));
    if (sel_per_sm < 1) return sc_fail(h, SC_ECUDA, "select kernel does not fit an SM");
    h->sel_grid_cap = nsm * sel_per_sm;
    return SC_OK;
}

extern "C" int sc_destroy(sc_handle h) {
    if (!h) return SC_EINVAL;
    if (h->d_tables || h->d_stats) {     // (a handle whose model was refused never touched CUDA)
        DeviceGuard g(h->device);
        cudaFree(h->d_tables); cudaFree(h->d_tfeat); cudaFree(h->d_sel); cudaFree(h->d_iota); cudaFree(h->d_stats);
    }
    delete h;
    return SC_OK;
}

extern "C" const char *sc_last_error(sc_handle h) { return h ? h->err : "null handle"; }
extern "C" int sc_contact_stats(sc_handle h, int64_t *out3, int reset) {
    if (!h || !out3) return SC_EINVAL;
    SC_ON_DEVICE(h);
    unsigned long long v[3];
    SC_CUDA(h, cudaMemcpy(v, h->d_stats, sizeof v, cudaMemcpyDeviceToHost));      // synchronises: a diagnostic, not a hot-path call
    out3[0] = (int64_t)v[0]; out3[1] = (int64_t)v[1]; out3[2] = (int64_t)v[2];
    if (reset) SC_CUDA(h, cudaMemset(h->d_stats, 0, sizeof v));
    return SC_OK;
}
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
    cudaMemcpy(out, h->d_prof, (SC_NPHASE + SC_NTRACE) * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemset(h->d_prof, 0, (SC_NPHASE + SC_NTRACE) * sizeof(long long));
    return SC_OK;
}
#endif

// Per-handle scratch grows in stream order: the old buffer is released after the work already queued on `st`, the new
// one exists before the work that follows -- no device-wide synchronisation, usable under stream capture.  (A handle
// is used from one stream at a time, include/samcon_b200.h.)
template <class T>
static int sc_reserve(sc_context *h, T **p, size_t *cap, size_t need, cudaStream_t st) {
    if (*cap >= need) return SC_OK;
    if (*p) SC_CUDA(h, cudaFreeAsync(*p, st));
    *p = nullptr; *cap = 0;
    need += need / 4;      // head room: a slowly growing batch does not reallocate every call
    SC_CUDA(h, cudaMallocAsync(p, need * sizeof(T), st));
    *cap = need;
    return SC_OK;
}

static int sc_launch_rollout(sc_context *h, const RolloutArgs &a_, cudaStream_t st) {
    RolloutArgs a = a_;
    const int ntiles = (a.S + TILE - 1) / TILE;
    // all SMs even when there are few tiles (walk.yml's own nSample = 2000 is 500 tiles: 3-4 warps on each of 148 SMs
    // beat 10 warps on 50 of them); the kernel deals the tiles evenly
    const int grid = ntiles < h->grid_cap ? ntiles : h->grid_cap;
    // warps per CTA: the fewest that keep the number of rounds the full CTA would need.  10 000 samples are 17 tiles on the
    // fullest SM = 2 rounds either way, and 9 warps do them without a tenth warp idling at every phase barrier (27.5 vs
    // 28.0 ms per window); walk.yml's own 2000 samples are 4 tiles per SM: 4 warps instead of 10.
    const int mine = (ntiles + grid - 1) / grid, rounds = (mine + WPB - 1) / WPB;
    const int nw = (mine + rounds - 1) / rounds;
    // helper warps (no tile, assembly calls only) up to what the register file holds next to the tile warps; none for the
    // launches that only evaluate features or costs
    static const int env_helpers = getenv("SC_HELPERS") ? atoi(getenv("SC_HELPERS")) : -1;      // developer knob
    // (measured: 4 % at walk.yml's own 2000 samples = 4 tile warps per CTA, nothing at 9-10 tile warps)
    int nhelp = a.nsteps > 0 ? (env_helpers >= 0 ? env_helpers : (nw <= WPB / 2 ? HELPER_WARPS : 0)) : 0;
    if (nhelp > HELPER_WARPS + WPB - nw) nhelp = HELPER_WARPS + WPB - nw;
    const int threads = 32 * (nw + nhelp);
    const int smem = (TABLE_WORDS + nw * WORDS_PER_TILE + COOP_WORDS) * 4;
    a.tile_warps = nw;
    // Phase barriers: all five keep the warps in the same code (instruction fetch), but the one behind the shared assembly
    // already aligns them once per sub-step, and a CTA of exactly 9 tile warps (5000 or 10 000 samples per GPU) is measured
    // 2 % faster with only that one and the one in front of the integration (mask 24); with 4 or 10 warps all five are as
    // good or better (interleaved A/B, profiles/r3_barrier_masks.log)
    static const int env_mask = getenv("SC_BARRIERS") ? atoi(getenv("SC_BARRIERS")) : -1;        // developer knob: 24 or 31
    const bool few = env_mask >= 0 ? env_mask == 24 : nw == 9;
    const bool box = h->h_tables.self.nbox > 0;
#if defined(SC_PROFILE)
    RolloutArgs ap = a;
    if (!h->d_prof) {
        cudaMalloc(&h->d_prof, (SC_NPHASE + SC_NTRACE) * sizeof(long long));
        cudaMemset(h->d_prof, 0, (SC_NPHASE + SC_NTRACE) * sizeof(long long));
        long long *tr = h->d_prof + SC_NPHASE;
        cudaMemcpyToSymbol(sc::sc_trace, &tr, sizeof tr);
    }
    ap.prof = a.nsteps > 0 ? h->d_prof : nullptr;
    const RolloutArgs &al = ap;
#else
    const RolloutArgs &al = a;
#endif
    if (box && few) sc_rollout_kernel<WPB, TablesBox, 24><<<grid, threads, smem, st>>>(h->d_tables, al);
    else if (box) sc_rollout_kernel<WPB, TablesBox, 31><<<grid, threads, smem, st>>>(h->d_tables, al);
    else if (few) sc_rollout_kernel<WPB, TablesCapsule, 24><<<grid, threads, smem, st>>>(h->d_tables, al);
    else sc_rollout_kernel<WPB, TablesCapsule, 31><<<grid, threads, smem, st>>>(h->d_tables, al);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

// features of `n` states (rows of `states`) -> d_tfeat[n,18]
static int sc_features(sc_context *h, const float *states, int n, cudaStream_t st) {
    int rc = sc_reserve(h, &h->d_tfeat, &h->tfeat_cap, (size_t)n * 18, st);
    if (rc) return rc;
    RolloutArgs f = {};
    f.parents = states; f.children = 1; f.S = n; f.nsteps = 0; f.out_feat = h->d_tfeat;
    return sc_launch_rollout(h, f, st);
}

extern "C" int sc_window(sc_handle h, const float *parents, const float *parent_cache, int E, const int32_t *parent_idx,
                         int children, const float *actions, const float *targets, int nseq, const int32_t *sample_seq, int S,
                         int nsteps, float *out_state, float *out_cost, float *out_info, float *out_cache, int32_t *out_overflow,
                         void *stream) {
    if (!h) return SC_EINVAL;
    if (!parents || !actions || !targets || S < 1 || E < 1 || nseq < 1 || nsteps < 0)
        return sc_fail(h, SC_EINVAL, "sc_window: bad arguments (S=%d E=%d nseq=%d nsteps=%d)", S, E, nseq, nsteps);
    if (!parent_idx && (children < 1 || (long long)E * children < S))
        return sc_fail(h, SC_EINVAL, "sc_window: %d parents x %d children cannot seed %d samples", E, children, S);
    if (nseq > 1 && !sample_seq) return sc_fail(h, SC_EINVAL, "sc_window: nseq=%d needs sample_seq", nseq);
    cudaStream_t st = (cudaStream_t)stream;
    SC_ON_DEVICE(h);
    int rc = sc_features(h, targets, nseq, st);
    if (rc) return rc;
    RolloutArgs a = {};
    a.parent_cache = parent_cache; a.out_cache = out_cache; a.out_overflow = out_overflow; a.stats = h->d_stats;
    a.parents = parents; a.parent_idx = parent_idx; a.children = children > 0 ? children : 1; a.actions = actions;
    a.targets = targets; a.tfeat = h->d_tfeat; a.sample_seq = sample_seq; a.S = S; a.nsteps = nsteps;
    a.out_state = out_state; a.out_cost = out_cost; a.out_info = out_info;
    return sc_launch_rollout(h, a, st);
}

extern "C" int sc_step(sc_handle h, float *state, float *cache, const float *actions, int S, int nsteps, void *stream) {
    if (!h) return SC_EINVAL;
    if (!state || !actions || S < 1 || nsteps < 0) return sc_fail(h, SC_EINVAL, "sc_step: bad arguments");
    SC_ON_DEVICE(h);
    RolloutArgs a = {};
    a.parent_cache = cache; a.out_cache = cache;
    a.parents = state; a.children = 1; a.actions = actions; a.S = S; a.nsteps = nsteps; a.out_state = state;
    return sc_launch_rollout(h, a, (cudaStream_t)stream);
}

extern "C" int sc_cost(sc_handle h, const float *sim, const float *kin, int S, float *out, void *stream) {
    if (!h) return SC_EINVAL;
    if (!sim || !kin || !out || S < 1) return sc_fail(h, SC_EINVAL, "sc_cost: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SC_ON_DEVICE(h);
    int rc = sc_features(h, kin, S, st);
    if (rc) return rc;
    rc = sc_reserve(h, &h->d_iota, &h->iota_cap, (size_t)S, st);
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

extern "C" int sc_sample_gen(sc_handle h, const float *targets, int nseq, const float *limits, const float *noise,
                             const int32_t *perm, int gaussian, uint64_t seed, uint64_t window, int S, float *actions,
                             void *stream) {
    if (!h) return SC_EINVAL;
    if (!targets || !limits || !actions || S < 1 || nseq < 1) return sc_fail(h, SC_EINVAL, "sc_sample_gen: bad arguments");
    SC_ON_DEVICE(h);
    const int nj = h->h_tables.nj;
    const long long n = (long long)nseq * S * nj;
    sc_sample_gen_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(targets, limits, noise, perm, gaussian, seed,
                                                                                         window, S, nseq, nj, actions);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_select(sc_handle h, const float *cost, const int32_t *seg_offsets, int nseg, int E, int32_t *elite_idx,
                         float *elite_cost, int total, void *stream) {
    if (!h) return SC_EINVAL;
    if (!cost || !seg_offsets || !elite_idx || nseg < 1 || E < 1 || total < 0)
        return sc_fail(h, SC_EINVAL, "sc_select: bad arguments");
    SC_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    int Epad = 2;
    while (Epad < E) Epad <<= 1;
    // chunks: one CTA each, all co-resident (grid-wide barriers).  total = seg_offsets[nseg] - seg_offsets[0] is passed by
    // the caller so that no device-to-host read is needed here.
    const int cap = h->sel_grid_cap;
    if (nseg >= cap) return sc_fail(h, SC_EINVAL, "sc_select: %d segments exceed the %d co-resident CTAs; split the call", nseg, cap);
    int chunk = (int)(((long long)total + (cap - nseg) - 1) / (cap - nseg));
    if (chunk < 2048) chunk = 2048;
    long long grid = (long long)total / chunk + nseg;          // >= sum over segments of ceil(n / chunk)
    if (grid > cap) grid = cap;
    if (Epad > SEL_TILE) grid = cap;      // the grid-wide sort uses every CTA
    const size_t nkeys = (size_t)nseg * Epad, nhist = (size_t)8 * nseg * 256 + nseg;
    int rc = sc_reserve(h, &h->d_sel, &h->sel_cap, nkeys + (nhist + 1) / 2, st);
    if (rc) return rc;
    SelectArgs a;
    a.cost = cost; a.seg = seg_offsets; a.nseg = nseg; a.E = E; a.Epad = Epad; a.chunk = chunk;
    a.buf = h->d_sel; a.ghist = reinterpret_cast<unsigned int *>(h->d_sel + nkeys); a.gcount = a.ghist + (size_t)8 * nseg * 256;
    a.elite_idx = elite_idx; a.elite_cost = elite_cost;
    SC_CUDA(h, cudaMemsetAsync(a.ghist, 0, nhist * sizeof(unsigned int), st));
    void *args[] = {&a};
    SC_CUDA(h, cudaLaunchCooperativeKernel((void *)sc_select_kernel, dim3((unsigned)grid), dim3(SEL_THREADS), args,
                                           SEL_TILE * sizeof(unsigned long long), st));
    h->launches++;
    return SC_OK;
}

extern "C" int sc_gather_rows_peer(sc_handle h, const uint64_t *peer_base, const int32_t *src_peer, const int32_t *src_row, int n,
                                   int width, float *dst, void *stream) {
    if (!h) return SC_EINVAL;
    if (!peer_base || !src_peer || !src_row || !dst || n < 1 || width < 1) return sc_fail(h, SC_EINVAL, "sc_gather_rows_peer: bad arguments");
    SC_ON_DEVICE(h);
    const size_t tot = (size_t)n * width;
    sc_gather_rows_peer_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const unsigned long long *>(peer_base), src_peer, src_row, n, width, dst);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_measure_fp32_peak(int device, double *tflops) {
    if (!tflops) return SC_EINVAL;
    DeviceGuard g(device);
    if (!g.ok) return SC_ECUDA;
    int nsm = 0;
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return SC_ECUDA;
    const int blocks = nsm * 8, iters = 4096;
    float *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)) != cudaSuccess) return SC_ENOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {          // first repetitions warm the clocks up; the best one is the peak
        cudaEventRecord(e0);
        sc_ffma_peak_kernel<<<blocks, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return SC_ECUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (double)blocks * 256 * iters * 16 * 8 * 2 / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return SC_OK;
}

extern "C" int sc_motion_states(sc_handle h, const float *poses, const int32_t *frame_offsets, const double *frame_dt, int nseq,
                                const double *times, int nt, float *out, void *stream) {
    if (!h) return SC_EINVAL;
    if (!poses || !frame_offsets || !frame_dt || !times || !out || nseq < 1 || nt < 1)
        return sc_fail(h, SC_EINVAL, "sc_motion_states: bad arguments");
    SC_ON_DEVICE(h);
    const int nj = h->h_tables.nj;
    const long long n = (long long)nseq * nt * (nj + 1);
    sc_motion_states_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(poses, frame_offsets, frame_dt, times, nseq, nt,
                                                                                            nj, out);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

extern "C" int sc_gather_rows(sc_handle h, const float *src, const int32_t *idx, int n, int width, float *dst, void *stream) {
    if (!h) return SC_EINVAL;
    if (!src || !idx || !dst || n < 1 || width < 1) return sc_fail(h, SC_EINVAL, "sc_gather_rows: bad arguments");
    SC_ON_DEVICE(h);
    const size_t tot = (size_t)n * width;
    sc_gather_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, width, dst);
    h->launches++;
    SC_CUDA(h, cudaGetLastError());
    return SC_OK;
}

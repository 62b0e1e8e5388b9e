/*
 * samcon_b200.h -- C ABI of the B200 SamCon sample-rollout path.
 *
 * The reference has no FFI layer: its seam is the worker-pool protocol
 *   ["init"|"sim"|"close", data]          /root/reference/samcon/core/samcon.py:15-52
 * driven by SamCon.sample()                /root/reference/samcon/core/samcon.py:144-198
 * plus the env calls the workers make      /root/reference/envs/humanoid_tracking_env.py:178-200.
 * Each entry point below names the reference call(s) it replaces.  All data pointers are DEVICE
 * pointers to contiguous float32 / int32 arrays owned by the caller (PyTorch tensors on the Python
 * side); every call enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 * synchronising (sc_contact_stats excepted); scratch buffers grow with cudaMallocAsync on that stream.
 * Every call runs on the handle's device and restores the caller's current device before returning.
 * Return value: 0 on success, negative SC_E* otherwise; sc_last_error() gives text.
 * One handle per device and per host thread, used from ONE stream at a time (per-handle scratch is not
 * per-stream: calls on a second stream must be ordered after the first by the caller); handles are
 * not thread-safe.
 */
#ifndef SAMCON_B200_H
#define SAMCON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_STATE_DIM 132  /* /root/reference/envs/humanoid_stable_pd.py:137-165 */
#define SC_MAX_BODIES 18  /* moving bodies after folding fixed joints */
#define SC_MAX_LEVELS 8
#define SC_MAX_CAND 80    /* ground-contact candidate points */
#define SC_MAX_CPOINTS 24 /* cost points = URDF links */
#define SC_MAX_CONTACTS 8 /* per-sample cap, deepest first */
#define SC_CACHE_DIM (2 * SC_MAX_CONTACTS) /* contact-cache row: candidate ids (as floats, -1 = empty), normal impulses */
#define SC_MAX_PRIMS 24   /* self-collision primitives (sphere-swept segments) */
#define SC_MAX_PAIRS 128  /* self-collision primitive pairs */
#define SC_MAX_BOXES 8    /* box primitives with an exact narrow phase */

#define SC_OK 0
#define SC_EINVAL (-1)
#define SC_ECUDA (-2)
#define SC_ENOMEM (-3)

/* Host-side character + engine description, produced by samcon_b200.model.Character.device_tables()
 * (what pybullet.loadURDF + setPhysicsEngineParameter hold for the reference:
 *  /root/reference/envs/humanoid_stable_pd.py:9-98, /root/reference/envs/humanoid_tracking_env.py:43-50,75-76).
 * All pointers are HOST pointers, copied by sc_create. */
typedef struct sc_model {
    int32_t nb, nlev, ncand, ncp, nee;
    const int32_t *parent;       /* [nb]  parent body, -1 for the root; bodies are sorted by tree depth */
    const int32_t *jidx;         /* [nb]  position of the body's joint in the 132-state (root: -1) */
    const int32_t *level_start;  /* [nlev+1] */
    const int32_t *cand_body;    /* [ncand] */
    const int32_t *cp_body;      /* [ncp] */
    const int32_t *ee;           /* [nee] end effectors as cost-point indices */
    const float *jorg;           /* [nb,3] joint origin in the parent body frame */
    const float *inertia_dyn;    /* [nb,10] m, h=m*c (3), I about the body origin xx xy xz yy yz zz */
    const float *inertia_pd;     /* [nb,10] same, rigid-body model of the stable-PD controller */
    const float *kp, *kd, *maxf; /* [nb] per body (root entries ignored) */
    const float *cand;           /* [ncand,4] point in body frame (xyz) and sphere radius */
    const float *cp_off;         /* [ncp,3] link CoM in its body frame */
    const float *cp_mass;        /* [ncp] */
    const float *joint_weights;  /* [nb] pose-cost weight per body (root ignored) */
    float cost_weights[5];       /* pose root ee balance com */
    float height;
    float dt;                    /* 1/fps_sim */
    int32_t num_substeps, num_solver_iters, max_contacts;
    float gravity[3], friction, erp, contact_threshold, max_velocity;
    /* link-vs-link contact (URDF_USE_SELF_COLLISION | ..._EXCLUDE_ALL_PARENTS,
     * /root/reference/envs/humanoid_stable_pd.py:10-13); npair = 0 switches it off (cfg self_collision: False) */
    int32_t nprim, npair;
    const int32_t *prim_body;    /* [nprim] */
    const float *prim;           /* [nprim,7] sphere-swept segment in the body frame: p0 xyz, p1 xyz, radius */
    const int32_t *pair;         /* [npair,2] primitive indices; links neither equal nor ancestor-related */
    float friction_self;         /* link lateral friction x link lateral friction */
    float warmstart;             /* 0..1: factor on a persisting contact's cached normal impulse (0 = stateless contacts) */
    float spinning_friction, spinning_friction_self; /* must be 0: the torsional row exists in the CPU oracle only */
    /* exact boxes in the link-vs-link narrow phase (feet, upper arms, forearms): nbox = 0 treats every box as its capsule
     * stand-in `prim` describes */
    int32_t nbox;
    const int32_t *prim_box;     /* [nprim] box index of the primitive, -1 if it is a sphere / capsule */
    const float *box;            /* [nbox,15] body frame: centre xyz, axes R (row major, v_body = R v_box), half extents */
} sc_model;

typedef struct sc_context *sc_handle;

/* HumanoidTrackingEnv(cfg, viz=False) per worker -- samcon.py:19, humanoid_tracking_env.py:15-71 */
int sc_create(sc_handle *out, const sc_model *model, int device);
/* worker "close" -- samcon.py:47-49 */
int sc_destroy(sc_handle h);
const char *sc_last_error(sc_handle h);

/* SamCon.get_batch_action -- samcon.py:101-142, for nseq sequences at once (rows [q S, (q+1) S) belong to sequence q).
 * targets [nseq,132]; limits [3*nj] (cfg joint_noises); noise [nseq*S,3*nj] in [0,1) (the reference's np.random.random
 * draw) or NULL to draw Philox-4x32-10 uniforms keyed by (seed + q, window) and counted from the sequence's first row;
 * perm [nseq*S] row permutation inside each sequence (np.random.shuffle) or NULL; gaussian!=0 reads `noise` as standard
 * normals (or draws Box-Muller normals) scaled by sqrt(limits).  actions [nseq*S,4*nj] xyzw. */
int sc_sample_gen(sc_handle h, const float *targets, int nseq, const float *limits, const float *noise, const int32_t *perm,
                  int gaussian, uint64_t seed, uint64_t window, int S, float *actions, void *stream);

/* AmassMotionData.getSpecTimeState for every (sequence, time) pair -- data_loaders/amass_motion_data.py:48-109 with the
 * finite-difference velocities of utils/bullet_utils.py:9-45.  poses [sum of frames, 3+4*(nj+1)] = the sequences' (T,75)
 * pose arrays one after the other (root position, root quaternion, joint quaternions, xyzw); frame_offsets [nseq+1] first
 * row of each sequence; frame_dt [nseq] key-frame duration 1/fr (double); times [nt] query times in seconds (double).
 * out [nseq, nt, 132] target states (cycling past the last frame with the cycle offset, like the reference). */
int sc_motion_states(sc_handle h, const float *poses, const int32_t *frame_offsets, const double *frame_dt, int nseq,
                     const double *times, int nt, float *out, void *stream);

/* Worker "sim" for all samples of all sequences at once -- samcon.py:28-41 + env.step + cost
 * (humanoid_tracking_env.py:184-200, 78-176).  Sample k restores parents[p] with p = parent_idx ? parent_idx[k] : k/children
 * (restore_pb_state, humanoid_tracking_env.py:181-182: the .bullet snapshot carries the contact cache, so parent_cache[p]
 * [E,SC_CACHE_DIM] is restored with it; NULL = empty caches, the first window), holds actions[k] for nsteps x {stable PD;
 * stepSimulation}, and is scored against targets[sample_seq ? sample_seq[k] : 0].  Outputs in sample order: out_state
 * [S,132], out_cost [S], out_info [S,5] = pose, root, ee, balance, com, out_cache [S,SC_CACHE_DIM] (save_pb_state,
 * humanoid_tracking_env.py:178-179), out_overflow [S] = in how many of the window's sub-steps the sample had more than
 * max_contacts candidate points inside the margin (low 16 bits) / touching (high 16 bits), see sc_contact_stats; any output
 * may be NULL. */
int sc_window(sc_handle h, const float *parents, const float *parent_cache, int E, const int32_t *parent_idx, int children,
              const float *actions, const float *targets, int nseq, const int32_t *sample_seq, int S, int nsteps,
              float *out_state, float *out_cost, float *out_info, float *out_cache, int32_t *out_overflow, void *stream);

/* env.step without the cost: used by the replay facade (scripts/eval_samcon.py:67-81). state [S,132] and cache
 * [S,SC_CACHE_DIM] (or NULL: empty caches in, result dropped) are updated in place. */
int sc_step(sc_handle h, float *state, float *cache, const float *actions, int S, int nsteps, void *stream);

/* Diagnostic, SYNCHRONISES the device.  out3[0] = sample-sub-steps in which more candidate points were inside the contact
 * margin than max_contacts (the deepest were kept: what is dropped first are speculative points that do not touch yet),
 * out3[1] = sample-sub-steps simulated by sc_window / sc_step since the handle was created or last reset, out3[2] = sample-
 * sub-steps in which more than max_contacts points were actually TOUCHING (distance <= 0), i.e. a live contact was dropped. */
int sc_contact_stats(sc_handle h, int64_t *out3, int reset);

/* compute_total_cost for explicit (sim, kin) pairs -- humanoid_tracking_env.py:78-90. out [S,6] total + 5 terms */
int sc_cost(sc_handle h, const float *sim, const float *kin, int S, float *out, void *stream);

/* Trajectory.select('greedy') -- /root/reference/samcon/core/trajectory.py:108-112, segmented: for each of nseg
 * segments [seg_offsets[i], seg_offsets[i+1]) writes the indices (global, into cost[]) of its E smallest costs
 * in ascending (cost, index) order to elite_idx[i*E ...] (a segment shorter than E: its entries, then -1 / +inf).  NaN costs
 * sort last.  Scratch is owned by the handle and grows in stream order.
 * total = seg_offsets[nseg] - seg_offsets[0], the number of costs the call looks at (the offsets live on the device; the
 * host needs the figure to size the launch).  One cooperative launch: a segment is split over as many CTAs as its length
 * needs, so a 10^6-sample window uses the whole GPU and 64 sequences are selected side by side. */
int sc_select(sc_handle h, const float *cost, const int32_t *seg_offsets, int nseg, int E, int32_t *elite_idx,
              float *elite_cost, int total, void *stream);

/* restore_pb_state/save_pb_state replacement -- humanoid_tracking_env.py:178-182: rows of src [*,width] picked by
 * idx [n] are copied to dst [n,width] (elite states and contact caches become next window's parents, all in HBM);
 * idx -1 (sc_select's padding) gives a zero row. */
int sc_gather_rows(sc_handle h, const float *src, const int32_t *idx, int n, int width, float *dst, void *stream);

/* The multi-GPU form of the same restore: row r of dst [n,width] = row src_row[r] of the [*,width] array at device address
 * peer_base[src_peer[r]].  peer_base [npeer] holds addresses valid in THIS process -- the other ranks' result buffers
 * mapped through CUDA VMM / symmetric memory -- so the elites' snapshots cross NVLink inside one gather kernel instead of
 * a collective (the reference's master gathers every worker's pickled results, samcon.py:172-176). */
int sc_gather_rows_peer(sc_handle h, const uint64_t *peer_base, const int32_t *src_peer, const int32_t *src_row, int n,
                        int width, float *dst, void *stream);

/* Measurement aid, SYNCHRONISES: FP32 FFMA throughput of `device` in TFLOP/s (a register-only FFMA loop on every SM, best of
 * five), the measured counterpart of the nominal 148 x 128 x 2 x clock figure bench.py's roofline is quoted against. */
int sc_measure_fp32_peak(int device, double *tflops);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t sc_launch_count(sc_handle h);

/* shared-memory bytes and samples per CTA the rollout kernel was configured with (for DESIGN/bench reporting) */
int sc_rollout_config(sc_handle h, int *smem_bytes, int *samples_per_cta, int *threads_per_cta);

#ifdef __cplusplus
}
#endif
#endif

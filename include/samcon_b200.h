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
 * synchronising.  Return value: 0 on success, negative SC_E* otherwise; sc_last_error() gives text.
 * One handle per device and per host thread; handles are not thread-safe.
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
#define SC_MAX_PRIMS 24   /* self-collision primitives (sphere-swept segments) */
#define SC_MAX_PAIRS 128  /* self-collision primitive pairs */

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
} sc_model;

typedef struct sc_context *sc_handle;

/* HumanoidTrackingEnv(cfg, viz=False) per worker -- samcon.py:19, humanoid_tracking_env.py:15-71 */
int sc_create(sc_handle *out, const sc_model *model, int device);
/* worker "close" -- samcon.py:47-49 */
int sc_destroy(sc_handle h);
const char *sc_last_error(sc_handle h);

/* SamCon.get_batch_action -- samcon.py:101-142.  target [132]; limits [3*nj] (cfg joint_noises);
 * noise [S,3*nj] in [0,1) (the reference's np.random.random draw) or NULL to draw Philox-4x32-10 uniforms
 * from (seed, window); perm [S] row permutation (np.random.shuffle) or NULL; gaussian!=0 reads `noise`
 * as standard normals (or draws Box-Muller normals) scaled by sqrt(limits).  actions [S,4*nj] xyzw. */
int sc_sample_gen(sc_handle h, const float *target, const float *limits, const float *noise, const int32_t *perm,
                  int gaussian, uint64_t seed, uint64_t window, int S, float *actions, void *stream);

/* Worker "sim" for all samples of all sequences at once -- samcon.py:28-41 + env.step + cost
 * (humanoid_tracking_env.py:184-200, 78-176).  Sample k restores parents[parent_idx ? parent_idx[k] : k/children],
 * holds actions[k] for nsteps x {stable PD; stepSimulation}, and is scored against
 * targets[sample_seq ? sample_seq[k] : 0].  Outputs in sample order: out_state [S,132], out_cost [S],
 * out_info [S,5] = pose, root, ee, balance, com. */
int sc_window(sc_handle h, const float *parents, int E, const int32_t *parent_idx, int children,
              const float *actions, const float *targets, int nseq, const int32_t *sample_seq, int S, int nsteps,
              float *out_state, float *out_cost, float *out_info, void *stream);

/* env.step without the cost: used by the replay facade (scripts/eval_samcon.py:67-81). state [S,132] in place. */
int sc_step(sc_handle h, float *state, const float *actions, int S, int nsteps, void *stream);

/* compute_total_cost for explicit (sim, kin) pairs -- humanoid_tracking_env.py:78-90. out [S,6] total + 5 terms */
int sc_cost(sc_handle h, const float *sim, const float *kin, int S, float *out, void *stream);

/* Trajectory.select('greedy') -- /root/reference/samcon/core/trajectory.py:108-112, segmented: for each of nseg
 * segments [seg_offsets[i], seg_offsets[i+1]) writes the indices (global, into cost[]) of its E smallest costs
 * in ascending (cost, index) order to elite_idx[i*E ...].  NaN costs sort last.  scratch: see sc_select_scratch. */
int sc_select(sc_handle h, const float *cost, const int32_t *seg_offsets, int nseg, int E, int32_t *elite_idx,
              float *elite_cost, void *stream);

/* restore_pb_state/save_pb_state replacement -- humanoid_tracking_env.py:178-182: rows of src [*,width] picked by
 * idx [n] are copied to dst [n,width] (elite states become next window's parents, all in HBM). */
int sc_gather_rows(sc_handle h, const float *src, const int32_t *idx, int n, int width, float *dst, void *stream);

/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t sc_launch_count(sc_handle h);

/* shared-memory bytes and samples per CTA the rollout kernel was configured with (for DESIGN/bench reporting) */
int sc_rollout_config(sc_handle h, int *smem_bytes, int *samples_per_cta, int *threads_per_cta);

#ifdef __cplusplus
}
#endif
#endif

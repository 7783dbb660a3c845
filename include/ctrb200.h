/*
 * ctrb200.h -- C ABI of libctrb200.so, the B200-native (sm_100a) hot path of
 * CTR DaPP (ConorMesser/ctr-design-and-path-plan): batched strain-based
 * concentric-tube model + tube-vs-environment collision/distance + cost.
 *
 * Each entry point names the reference interface it replaces (file:line in the
 * upstream repository).  The reference is pure Python over python-fcl/SciPy,
 * so "the reference's FFI for this path" is a ctypes binding; INTEGRATION.md
 * shows the stub a maintainer would add to ctrdapp/model and ctrdapp/collision.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types; nothing throws.
 *   - every function returns CTRB_OK (0) or a negative ctrb_status;
 *     ctrb_last_error() gives a thread-local message for the last failure.
 *   - `mem` says where ALL data pointers of a batch call live:
 *     CTRB_MEM_HOST (the library stages through pinned buffers and copies
 *     H2D/D2H itself, synchronous on return) or CTRB_MEM_DEVICE (pointers are
 *     device pointers on the ctx's device; the call is stream-ordered on the
 *     ctx's stream and returns without synchronising).  The exceptions are
 *     named where they occur: entry points that return a host scalar
 *     (ctrb_argmin_cost, the *_last_* / counter / status queries),
 *     ctrb_check_collision_batch without CTRB_OPT_MAX_NODES_HINT, and the
 *     first call at a larger batch size, when a grow-only scratch buffer is
 *     reallocated.  Indices in device memory are not validated: the caller
 *     guarantees 0 <= idx[b][t] < npts[t].
 *   - handles (model, env) are immutable after create: concurrent batch calls
 *     with different ctx objects are legal.  A ctx owns one CUDA stream and
 *     its scratch buffers and must not be used from two threads at once.
 *   - SE(3) matrices are 4x4 row-major, position in [0:3,3] (kinematic.py,
 *     collision_checker.py:124-126); twists are [omega; v].
 *   - collision sentinel: obstacle_min / goal_dist == -1.0 exactly on contact
 *     (collision_checker.py:70; callers test `< 0`, rrt.py:116,120); DBL_MAX
 *     when there is no obstacle / no goal.
 */
#ifndef CTRB200_H
#define CTRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTRB_MAX_TUBES 4
#define CTRB_MAX_DOF 8

typedef enum {
    CTRB_OK = 0,
    CTRB_ERR_INVALID = -1,      /* bad argument (ValueError in the reference) */
    CTRB_ERR_CUDA = -2,         /* CUDA runtime failure; no CPU fallback exists */
    CTRB_ERR_UNSUPPORTED = -3,  /* NotImplementedError in the reference */
    CTRB_ERR_NO_DEVICE = -4,
    CTRB_ERR_ALLOC = -5
} ctrb_status;

typedef enum { CTRB_MEM_HOST = 0, CTRB_MEM_DEVICE = 1 } ctrb_mem;
typedef enum { CTRB_F64 = 64, CTRB_F32 = 32 } ctrb_precision;

/* strain_bases.py:227-268 get_strains() names */
typedef enum {
    CTRB_BASE_CONSTANT = 0,
    CTRB_BASE_LINEAR = 1,
    CTRB_BASE_QUADRATIC = 2,
    CTRB_BASE_PURE_HELIX = 3,
    CTRB_BASE_LINEAR_HELIX = 4,
    CTRB_BASE_TORSION_HELIX = 5,
    CTRB_BASE_TORSION_LINEAR_HELIX = 6,
    CTRB_BASE_FULL = 7
} ctrb_base_kind;

typedef enum { CTRB_MODEL_KINEMATIC = 0, CTRB_MODEL_STATIC = 1 } ctrb_model_type;

/* What create_model(configuration, q) reads (model_factory.py:12-78). */
typedef struct {
    int32_t model_type;                 /* ctrb_model_type */
    int32_t tube_num;                   /* 'tube_number' */
    int32_t base_kind[CTRB_MAX_TUBES];  /* 'strain_bases' */
    int32_t q_dof[CTRB_MAX_TUBES];      /* 'q_dof' */
    double q[CTRB_MAX_TUBES][CTRB_MAX_DOF];
    double tube_length[CTRB_MAX_TUBES]; /* 'tube_lengths' */
    double delta_x;                     /* 'delta_x' */
    double radius_outer[CTRB_MAX_TUBES];/* 'tube_radius'.outer (static stiffness, static.py:561-585) */
    double radius_inner[CTRB_MAX_TUBES];
    int32_t static_degree;              /* 'degree' (static_bases.py) */
    int32_t static_basis_type;          /* 0 = last_strain_base (the only one that runs) */
} ctrb_model_desc;

typedef enum {
    CTRB_SHAPE_NONE = 0,
    CTRB_SHAPE_SPHERE = 1,
    CTRB_SHAPE_BOX = 2,
    CTRB_SHAPE_CYLINDER = 3,
    CTRB_SHAPE_MESH = 4
} ctrb_shape;

/* One primitive obstacle or the goal (init_collision.py:128-142). */
typedef struct {
    int32_t shape;
    double center[3];   /* 'transform' */
    double rot[9];      /* row-major; rotate_align('direction') for Cylinder, identity otherwise */
    double dims[3];     /* Sphere: radius ; Box: x,y,z lengths ; Cylinder: radius, length */
} ctrb_prim;

/* cost selector (heuristic_factory.py:173-203) */
typedef enum {
    CTRB_COST_SOAPWG = 0,           /* square_obstacle_avg_plus_weighted_goal */
    CTRB_COST_ONLY_GOAL = 1,        /* only_goal_distance */
    CTRB_COST_FTL = 2,              /* follow_the_leader */
    CTRB_COST_FTL_W_INSERTION = 3,  /* follow_the_leader_w_insertion */
    CTRB_COST_FTL_TRANSLATION = 4   /* follow_the_leader_translation */
} ctrb_cost_type;

typedef struct {
    int32_t cost_type;
    double goal_weight;       /* SOAPWG */
    int32_t only_tip;         /* FTL family */
    double insertion_weight;  /* FTL_W_INSERTION */
} ctrb_cost_desc;

typedef struct ctrb_model ctrb_model;
typedef struct ctrb_env ctrb_env;
typedef struct ctrb_ctx ctrb_ctx;

/* ---- library ---- */
const char* ctrb_last_error(void);
const char* ctrb_version(void);
int ctrb_device_count(void);

/* ---- execution context: device, stream, scratch ---- */
int ctrb_ctx_create(int device, ctrb_ctx** out);
void ctrb_ctx_destroy(ctrb_ctx* ctx);
int ctrb_ctx_synchronize(ctrb_ctx* ctx);
/* Launch on a caller-owned stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream)
 * so that the caller's events and collectives order with the kernels; NULL restores the ctx's own. */
int ctrb_ctx_set_stream(ctrb_ctx* ctx, void* cuda_stream);
/* kernels launched by this ctx so far (bench.py's gpu_launches). */
int64_t ctrb_ctx_launch_count(const ctrb_ctx* ctx);
/* milliseconds the ctx's stream spent between the two most recent
 * ctrb_ctx_mark() calls (CUDA events on the launching stream). */
int ctrb_ctx_mark(ctrb_ctx* ctx, int which /*0 = start, 1 = stop*/);
int ctrb_ctx_elapsed_ms(ctrb_ctx* ctx, float* ms);

/* Device duration of the most recent launch of one hot kernel on this ctx (CUDA events recorded
 * on the launching stream around that launch; synchronises on the stop event). */
typedef enum {
    CTRB_KERNEL_CHAIN_EVAL = 0,   /* chain_eval_kernel (fused or given curves) */
    CTRB_KERNEL_STATIC_SOLVE = 1, /* static_solve_kernel */
    CTRB_KERNEL_GCURVE = 2,       /* kin_gcurve_kernel */
    CTRB_KERNEL_ETA_FTL = 3,      /* kin_eta_ftl_kernel */
    CTRB_KERNEL_STATIC_FAST = 4,  /* static_fast_kernel alone (STATIC_SOLVE spans it, the QR kernel and the curves) */
    CTRB_KERNEL_STATIC_QR = 5,    /* static_solve_kernel alone (the configurations the fast kernel handed back) */
    CTRB_KERNEL_STATIC_CURVES = 6 /* static_curves_kernel alone (integrate_g of the solved unknowns) */
} ctrb_kernel_id;
int ctrb_ctx_last_kernel_ms(ctrb_ctx* ctx, int kernel_id, float* ms);

/* Run-time switches of a ctx, for parity tests and measurement; the defaults are the product path.  A ctx is used
 * by one thread at a time, so an option set before a batch call holds for that call; nothing is read from the
 * process environment. */
typedef enum {
    CTRB_OPT_STATIC_SOLVER = 0,     /* 0 (default): explicit-inverse kernel, QR kernel for what it hands back;
                                       1: every configuration through the QR kernel (MINPACK's own factor machinery) */
    CTRB_OPT_STATIC_ONE_STEP = 1,   /* one-step sections (two nodes; rank-deficient Jacobian):
                                       0 (default): hybrd on the reduced system in the explicit-inverse kernel (the x^2
                                       curvature coefficient of such a section keeps its initial value);
                                       1: MINPACK's literal treatment in the QR kernel (null-space component decided by rounding) */
    CTRB_OPT_FAST_GENERIC_N = 2,    /* 1: run-time-n instantiation of the explicit-inverse kernel (default 0: size-templated) */
    CTRB_OPT_FAST_BLOCK_INVERT = 3, /* 0: general Gauss-Jordan inverse (default 1: block-structured for three tubes) */
    CTRB_OPT_EVAL_GROUP = 4,        /* 0: one warp per configuration in the collision kernel (default 1: grouped) */
    CTRB_OPT_GCURVE_BLOCKS = 5,     /* resident blocks per SM of the g-curve kernel's persistent grid (default 3) */
    CTRB_OPT_MAX_NODES_HINT = 6,    /* ctrb_check_collision_batch(CTRB_MEM_DEVICE): upper bound on the nodes of one
                                       configuration's curves; 0 (default) = read the offsets back, which synchronises */
    CTRB_OPT_STATIC_CURVES_SCAN = 7,/* ctrb_static_eval_batch: 1 = node positions from the warp-per-configuration scan kernel
                                       instead of the thread-per-configuration one (default 0); parity test of the two */
    CTRB_OPT_COUNT = 8
} ctrb_option;
int ctrb_ctx_set_option(ctrb_ctx* ctx, int option, int value);
int ctrb_ctx_get_option(const ctrb_ctx* ctx, int option, int* value);

/* Measured CUDA-core FMA peak of this device (dependent-free FMA chains on every SM), the
 * denominator of the compute roofline of the collision kernel: precision CTRB_F32 / CTRB_F64. */
int ctrb_measure_fma_peak(ctrb_ctx* ctx, int precision, double* tflops);

/* ---- model: replaces create_model() / Model.__init__ (model_factory.py:12-78, model.py:33-38) ---- */
int ctrb_model_create(ctrb_ctx* ctx, const ctrb_model_desc* desc, ctrb_model** out);
void ctrb_model_destroy(ctrb_model* m);
/* num_discrete_points (model.py:37-38) */
int ctrb_model_num_points(const ctrb_model* m, int32_t* npts /*[tube_num]*/);

/* ---- environment: replaces CollisionChecker.__init__ / add_obstacles / add_goal
 *      (collision_checker.py:31-33, init_collision.py:14-146).  Triangles are
 *      [n][3][3] doubles, already translated (float32 STL vertex promoted to
 *      double + 'transform', init_collision.py:82-86,125). ---- */
int ctrb_env_create(ctrb_ctx* ctx, const double* obstacle_tris, int32_t n_obstacle_tris, const ctrb_prim* prims,
                    int32_t n_prims, const ctrb_prim* goal /* may be NULL */, const double* goal_tris,
                    int32_t n_goal_tris, ctrb_env** out);
void ctrb_env_destroy(ctrb_env* e);
/* BVH statistics for reporting: nodes, leaves, max depth, triangles */
int ctrb_env_stats(const ctrb_env* e, int32_t out[4]);

/* ---- index rounding: replaces calculate_indices() (kinematic.py:201-274), batched.
 *      prev/next insertions are in the model's inverted coordinate (s = L - insertion). ---- */
int ctrb_calculate_indices_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const double* prev_ins /*[B][T]*/,
                                 const double* next_ins /*[B][T]*/, int32_t* prev_idx /*[B][T]*/,
                                 int32_t* new_idx /*[B][T]*/, int mem);

/* ---- g-curves: replaces Kinematic.solve_g (kinematic.py:82-130) for B configurations.
 *      Output is ragged and packed: tube n of config b owns nodes
 *      [node_offsets[b*T+n], node_offsets[b*T+n+1]) of g_out, each node a 4x4
 *      row-major matrix of `precision`.  With full == 0 tube n has
 *      npts[n]-idx[b][n] nodes (kinematic.py:100-103), with full != 0 npts[n]. ---- */
int ctrb_g_node_offsets(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx /*[B][T]*/, int full,
                        int64_t* node_offsets /*[B*T+1]*/, int mem);
int ctrb_solve_g_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx /*[B][T]*/,
                       const double* theta /*[B][T]*/, int full, int precision,
                       const int64_t* node_offsets /*[B*T+1]*/, void* g_out, int mem);

/* ---- eta + follow-the-leader: replaces Kinematic.solve_eta (kinematic.py:132-198).
 *      prev_g is packed like ctrb_solve_g_batch output (fp64) with prev_offsets;
 *      tube n of config b must hold npts[n]-prev_idx[b][n] nodes.
 *      ftl_out (may be NULL) is packed with the same offsets, 6 doubles per node;
 *      ftl_mag[b][0] = magnitude at the tip node of the last tube,
 *      ftl_mag[b][1] = mean magnitude over all nodes (follow_the_leader.py:29-38,89-116),
 *      [2],[3] = the same with the translation-only magnitude (:173-189).
 *      Exactly one of new_idx / velocity is given: solve_integrate derives the velocity
 *      from the rounded indices, (prev_idx - new_idx) * delta_x (kinematic.py:71); solve_eta's
 *      own signature takes it as a float list (kinematic.py:132). ---- */
int ctrb_eta_ftl_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* prev_idx /*[B][T]*/,
                       const int32_t* new_idx /*[B][T] or NULL*/, const double* velocity /*[B][T] or NULL*/,
                       const double* delta_theta /*[B][T]*/,
                       const double* prev_g, const int64_t* prev_offsets /*[B*T+1]*/,
                       double* eta_out /*[B][T][6]*/, double* ftl_out /* nullable */, double* ftl_mag /*[B][4]*/,
                       int mem);

/* ---- static model: replaces Static.solve_g (static.py:71-135): MINPACK hybrd on the 15- / 30-unknown
 *      equilibrium + integrate_g of the section curves, per configuration.
 *      g_out is packed like ctrb_solve_g_batch (fp64, same node_offsets: tube n has npts[n]-idx[n] nodes).
 *      initial_guess: NULL (Static.general_init_guess), or [nq] shared (x0_per_config = 0), or [B][nq].
 *      xtol <= 0 selects SciPy's default 1.49012e-8.  solution_q [B][nq], nfev [B] (MINPACK's count; SciPy
 *      reports 2 more), info [B] (MINPACK info: 1 = converged -- what solution.success means) may be NULL. ---- */
int ctrb_static_num_unknowns(const ctrb_model* m, int32_t* nq);
int ctrb_static_initial_guess(const ctrb_model* m, double* x0 /*[nq], host*/);
/* configurations of the most recent static solve on this ctx that the explicit-inverse kernel handed to the
 * QR kernel (singular Jacobian: one-step / zero-length sections).  Synchronises. */
int ctrb_static_last_fallbacks(ctrb_ctx* ctx, uint64_t* count);
/* why: [0] zero-length (or, with CTRB_OPT_STATIC_ONE_STEP = 1, one-step) section, [1] non-finite residual at the
 * initial guess, [2] / [3] singular Jacobian at the first / a later evaluation, [4] non-finite trial point,
 * [5] singular Broyden update.  Synchronises. */
int ctrb_static_last_fallback_reasons(ctrb_ctx* ctx, uint64_t out[6]);
int ctrb_static_residual_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx /*[B][T]*/,
                               const double* theta /*[B][T]*/, const double* q /*[B][nq]*/, double* F /*[B][nq]*/,
                               int mem);
int ctrb_static_solve_g_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx /*[B][T]*/,
                              const double* theta /*[B][T]*/, const double* initial_guess, int x0_per_config,
                              double xtol, const int64_t* node_offsets /*[B*T+1]*/, double* g_out /* nullable */,
                              double* solution_q, int32_t* nfev, int32_t* info, int mem);
/* integrate_g alone (static.py:165-198 through the six calls of :110-130): the section curves of GIVEN unknowns
 * q [B][nq], no root find.  g_out / node_offsets as above. */
int ctrb_static_curves_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx /*[B][T]*/,
                             const double* theta /*[B][T]*/, const double* q /*[B][nq]*/,
                             const int64_t* node_offsets /*[B*T+1]*/, double* g_out, int mem);

/* ---- collision on given curves: replaces CollisionChecker.check_collision
 *      (collision_checker.py:35-70) for B curves.  g is packed fp64 4x4 nodes with
 *      node_offsets as above; rad[T] are the tube radii. ---- */
int ctrb_check_collision_batch(ctrb_ctx* ctx, const ctrb_env* env, int64_t B, int32_t tube_num, const double* g,
                               const int64_t* node_offsets /*[B*T+1]*/, const double* rad /*[T]*/,
                               double* obstacle_min /*[B]*/, double* goal_dist /*[B]*/, int mem);

/* ---- fused hot path: g-curve (never leaves the SM) + collision/distance + own cost:
 *      replaces the solve_g -> check_collision -> heuristic_factory.create chain of
 *      RRT._single_iteration (rrt.py:110-129) for B independent configurations.
 *      cost follows rrt.py:116-129: NaN for colliding configs (the planner discards
 *      them), goal_dist clamped at 0 when the tip is at the goal.
 *      collide[b] = 1 iff obstacle_min[b] < 0. ---- */
int ctrb_eval_batch(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cost_desc,
                    int64_t B, const int32_t* idx /*[B][T]*/, const double* theta /*[B][T]*/,
                    const double* rad /*[T], host pointer always*/, double* obstacle_min /*[B]*/,
                    double* goal_dist /*[B]*/, double* cost /*[B]*/, uint8_t* collide /*[B], nullable*/, int mem);
/* The same chain for a static model handle: static solve (g-curves to device scratch) -> collision ->
 * cost, plus the per-configuration MINPACK info (1 = converged), which the reference discards
 * (static.py:102-104) and SURVEY 7 makes part of the output contract. */
int ctrb_static_eval_batch(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cost_desc,
                           int64_t B, const int32_t* idx, const double* theta, const double* rad,
                           double* obstacle_min, double* goal_dist, double* cost, uint8_t* collide,
                           int32_t* info /*[B], nullable*/, int32_t* nfev /*[B], nullable*/, int mem);

/* work counters of the most recent ctrb_eval_batch / ctrb_check_collision_batch on this ctx
 * (bench publishes them): [0] BVH nodes visited, [1] fp32 point-triangle
 * prefilter tests, [2] exact fp64 cylinder-triangle tests, [3] configurations. Synchronises. */
int ctrb_ctx_work_counters(ctrb_ctx* ctx, uint64_t out[4]);
/* device-side error flags of the most recent ctrb_eta_ftl_batch on this ctx: bit 0 = a tube of prev_g held fewer
 * nodes than npts - prev_idx (IndexError in the reference).  Host-buffer calls report this as CTRB_ERR_INVALID
 * themselves; device-buffer calls do not synchronise, so their caller asks here.  Synchronises. */
int ctrb_ctx_device_status(ctrb_ctx* ctx, int32_t* flags);

/* ---- nearest neighbours of the planner's node store: replaces DynamicTree.nearest_neighbor /
 *      find_all_nearest_neighbor (dynamic_tree.py:66-120), i.e. pympnn.KDTree_topologies.k_nearest_neighbor,
 *      for Q queries at once.  points [N][dim] and queries [Q][dim] are interleaved
 *      [insertion_0, rotation_0, insertion_1, ...] (dynamic_tree.py:57-60), dim = 2 * tube_num <= 8;
 *      topology[d] = 1 (linear) or 2 (circular, period 2 pi) (dynamic_tree.py:53); weight[d] multiplies the
 *      per-dimension difference.  idx_out / dist_out [Q][k], ascending distance, ties by lower index;
 *      missing entries (N < k) are -1 / +inf.  k <= 32. ---- */
int ctrb_knn_batch(ctrb_ctx* ctx, int32_t dim, int64_t N, const double* points, int64_t Q, const double* queries,
                   const double* weight /*[dim], host*/, const int32_t* topology /*[dim], host*/, int32_t k,
                   int32_t* idx_out, double* dist_out, int mem);

/* best (lowest finite cost) configuration of a cost vector: packed key reduction on the
 * device; the multi-GPU driver all-reduces (min) the returned key over NCCL.
 * key = monotone fp64 bits of the cost (upper 64 - 24 bits) ... see DESIGN.md. */
int ctrb_argmin_cost(ctrb_ctx* ctx, int64_t B, const double* cost, int64_t index_base, double* best_cost,
                     int64_t* best_index, int mem);
/* The same reduction left on the device (cost and key_index are device pointers; stream-ordered, no synchronisation):
 * key_index[0] = order-preserving signed 64-bit image of the lowest finite cost (INT64_MAX if there is none; decode
 * with ctrb_key_to_cost), key_index[1] = index_base + its lowest index (INT64_MAX if none).  One process per GPU then
 * needs two 8-byte all-reduce(min) steps -- the key, then the index among the ranks that hold that key -- to agree on
 * the best node (ctrdapp_b200/dist.py: reduce_best). */
int ctrb_argmin_cost_device(ctrb_ctx* ctx, int64_t B, const double* cost, int64_t index_base, int64_t* key_index);
double ctrb_key_to_cost(int64_t key);
int64_t ctrb_cost_to_key(double cost);

#ifdef __cplusplus
}
#endif
#endif /* CTRB200_H */

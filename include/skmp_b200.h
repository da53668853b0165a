/*
 * skmp_b200.h -- C ABI of the B200-native collision-constraint engine for scikit-motionplan.
 *
 * This is the drop-in boundary.  Every entry point replaces one call the reference makes into
 * its two un-vendored native/numeric dependencies (tinyfk's pybind11 module and skrobot's SDF
 * callables).  Reference call sites are cited per function as  file:line  relative to the
 * reference checkout (HiroIshida/scikit-motionplan).
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes; no C++ / torch types cross the boundary.
 *   - every function returns 0 on success, non-zero on failure; skb_last_error() returns the
 *     message of the calling thread's last failure.  No exception crosses the ABI.
 *   - "dev" pointers are CUDA device pointers on the current device, written/read on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).  The library never
 *     allocates result memory for the device entry points: the caller owns all outputs.
 *   - "*_host" variants take host pointers (pageable or pinned) and run a chunked, multi-stream
 *     H2D -> kernel -> D2H pipeline internally; they return after the results are in host memory.
 *   - dtype selects the arithmetic AND storage type of q / outputs: SKB_F32 or SKB_F64.
 *   - layouts are the reference's: row-major C order, configuration index outermost
 *       q      (N, D)          D = n_joints + {0 FIXED, 3 PLANER [x,y,theta], 6 FLOATING [x,y,z,r,p,y]}
 *       points (N, F, T)       T = 3 IGNORE / 6 RPY / 7 XYZW          (skmp/kinematics.py:57)
 *       jac    (N, F, T, D)                                           (skmp/kinematics.py:61)
 *       values (N, M), jac (N, M, D) for constraints                  (skmp/constraint.py:341-342)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SKMP_B200_H
#define SKMP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SKB_API __attribute__((visibility("default")))
#else
#define SKB_API
#endif

typedef struct skb_model skb_model; /* kinematic tree + sticky joint state  (tinyfk.KinematicModel) */
typedef struct skb_sdf skb_sdf;     /* world signed-distance field: union of primitives / voxel grid  */
typedef struct skb_plan skb_plan;   /* (features, control joints, base type, rot type) compiled for one device */

enum { SKB_JOINT_FIXED = 0, SKB_JOINT_REVOLUTE = 1, SKB_JOINT_PRISMATIC = 2 };
enum { SKB_BASE_FIXED = 0, SKB_BASE_PLANER = 1, SKB_BASE_FLOATING = 2 }; /* tinyfk.BaseType     */
enum { SKB_ROT_IGNORE = 0, SKB_ROT_RPY = 1, SKB_ROT_XYZW = 2 };          /* tinyfk.RotationType */
enum { SKB_F32 = 0, SKB_F64 = 1 };
enum { SKB_MODE_ALL = 0, SKB_MODE_CLOSEST = 1 }; /* only_closest_feature (skmp/constraint.py:277,692) */
enum { SKB_COM_POINT = 0, SKB_COM_STABILITY = 1 }; /* solve_com_fk / COMStabilityConst (skmp/constraint.py:899-928) */

SKB_API int skb_version(void);
SKB_API const char *skb_last_error(void);
/* number of CUDA devices visible; fails (non-zero) when there is none */
SKB_API int skb_device_count(int32_t *out_count);

/* ---------------------------------------------------------------- kinematic model
 * replaces tinyfk.KinematicModel(urdf)            skmp/kinematics.py:135,188,257
 * Links are given topologically sorted (parent index < child index, root first, parent -1).
 * Link i carries the joint that connects it to its parent: type, origin (xyz, rpy in the
 * parent link frame) and axis (joint frame).  "Joint ids" in this ABI are child-link indices.
 */
SKB_API int skb_model_create(int32_t n_links, const int32_t *parent, const int32_t *joint_type,
                             const double *origin_xyzrpy /* n_links*6 */, const double *axis /* n_links*3 */,
                             skb_model **out);
/* replaces KinematicModel.add_new_link(name, parent_id, position, rpy)   skmp/kinematics.py:110,211,280 */
SKB_API int skb_model_add_link(skb_model *m, int32_t parent, const double xyz[3], const double rpy[3],
                               int32_t *out_link_id);
/* replaces KinematicModel.set_q(joint_ids, angles(+base pose), base_type)   skmp/kinematics.py:96
 * q has n_joints (+3 PLANER, +6 FLOATING) entries; the state is sticky for all later plans. */
SKB_API int skb_model_set_q(skb_model *m, const int32_t *joint_links, int32_t n_joints, const double *q,
                            int32_t base_type);
SKB_API int skb_model_get_state(const skb_model *m, double *out_angles /* n_links */, double out_base_pose[6]);
SKB_API int skb_model_num_links(const skb_model *m, int32_t *out);
SKB_API int skb_model_clone(const skb_model *m, skb_model **out); /* copy.deepcopy of a kin map: constraint.py:518,567 */
SKB_API void skb_model_destroy(skb_model *m);

/* ---------------------------------------------------------------- signed distance field
 * replaces the opaque `sdf: Callable[[ndarray (M,3)], ndarray (M,)]`   skmp/constraint.py:268
 * (skrobot Box/Cylinder .sdf, skrobot.sdf.UnionSDF, GridSDF).  A skb_sdf is the union (min) of
 * its primitives.  pos / rot (row-major 3x3, world-from-local) place the primitive in the world.
 */
SKB_API int skb_sdf_create(skb_sdf **out);
SKB_API int skb_sdf_add_box(skb_sdf *s, const double pos[3], const double rot[9], const double half_extents[3]);
SKB_API int skb_sdf_add_cylinder(skb_sdf *s, const double pos[3], const double rot[9], double radius,
                                 double half_height);
SKB_API int skb_sdf_add_sphere(skb_sdf *s, const double pos[3], double radius);
/* voxel grid, values[ix][iy][iz] in C order sampled at  lb + (ix,iy,iz)*spacing  (local frame);
 * trilinear inside, `fill` (finite) outside, gradient 0 outside.  host_values: SKB_F32 or SKB_F64. */
SKB_API int skb_sdf_add_grid(skb_sdf *s, const double pos[3], const double rot[9], const double lb[3],
                             const double spacing[3], const int32_t dims[3], double fill,
                             const void *host_values, int32_t values_dtype);
SKB_API int skb_sdf_num_prims(const skb_sdf *s, int32_t *out);
SKB_API void skb_sdf_destroy(skb_sdf *s);
/* sdf(points (M,3)) -> values (M,) [+ analytic gradients (M,3)]; replaces the sdf(...) calls at
 * skmp/constraint.py:314,329,353,368,397,401 */
SKB_API int skb_sdf_eval(const skb_sdf *s, int32_t dtype, const void *pts_dev, int64_t M, void *out_vals_dev,
                         void *out_grads_dev /* nullable */, void *stream);

/* ---------------------------------------------------------------- plan
 * Binds what tinyfk.solve_fk receives on every call (feature link ids, control joint ids,
 * base_type, rot_type; skmp/kinematics.py:48-55) to the model's *current* sticky state and
 * compiles it once: non-controlled joints and the base pose are folded into constants, the tree
 * is cut to the feature ancestors and uploaded.  A plan is a snapshot: after skb_model_set_q /
 * skb_model_add_link create a new plan (what reflect_skrobot_model does, kinematics.py:86-96).
 */
SKB_API int skb_plan_create(const skb_model *m, const int32_t *feature_links, int32_t n_features,
                            const int32_t *joint_links, int32_t n_joints, int32_t base_type, int32_t rot_type,
                            skb_plan **out);
SKB_API int skb_plan_dims(const skb_plan *p, int32_t *out_dim_cspace, int32_t *out_n_feature,
                          int32_t *out_dim_tspace);
/* colkin.get_radius_list()            skmp/constraint.py:311,356 */
SKB_API int skb_plan_set_radii(skb_plan *p, const double *radii /* n_features */);
/* check_sphere_id_pairs / check_sphere_pair_sqdists   skmp/constraint.py:743-744
 * pairs index into the plan's feature list (0..n_features-1), thresholds = (r1+r2)^2 */
SKB_API int skb_plan_set_pairs(skb_plan *p, const int32_t *pairs /* 2*P */, int32_t n_pairs,
                               const double *rsum_sq /* P */);
/* weights of the plan's features for the centre of mass: link masses at the inertial origins, force / g for
 * "action" links (COMStabilityConst.__init__, skmp/constraint.py:857-897) */
SKB_API int skb_plan_set_weights(skb_plan *p, const double *weights /* n_features */);
/* tuning knobs (0 = auto): spheres per staged chunk, staging buffers per warp, warps per CTA */
SKB_API int skb_plan_set_tuning(skb_plan *p, int32_t chunk_features, int32_t n_buffers, int32_t warps_per_cta,
                                int32_t ctas_per_sm);
/* Measures the launch configuration (configurations per warp tile, warps per CTA, fused / unfused row steps) of the step
 * kernel that serves skb_collfree (pass the SDF) or skb_pairwise (s = NULL) for this plan, dtype, mode and output set
 * (bit 0 values, bit 1 jacobians, bit 2 validity bytes, bit 3 the transposed layout of skb_collfree_csc), on up to 262 144 of the caller's sample configurations and on
 * scratch outputs of its own; synchronous, 30-130 ms.  The choice is stored under a content key of the plan's structure:
 * every later launch of a structurally equal plan (same robot / features / joints, any sticky state) reuses it, and
 * SKB_TUNE_CACHE=<file> persists it across processes.  Without it, calls of fewer than 2^20 configurations use an
 * issue-slot model; longer calls measure once themselves (SKB_S2_AUTOTUNE=0 disables that).  Results never depend on the
 * choice (bit-identical).  out_choice (nullable): {unfused, G, warps}, out_choice[0] = -1 when nothing was tunable. */
SKB_API int skb_plan_tune(const skb_plan *p, const skb_sdf *s /* NULL: pairwise */, int32_t dtype, const void *q_dev,
                          int64_t N, double margin, int32_t mode, int32_t outputs, void *stream, int32_t *out_choice);
/* the stored choice for (plan structure, dtype, mode, outputs): out_choice = {unfused, G, warps}, {-1, ..} when none */
SKB_API int skb_plan_get_tuned(const skb_plan *p, const skb_sdf *s /* NULL: pairwise */, int32_t dtype, int32_t mode,
                               int32_t outputs, int32_t *out_choice);
SKB_API void skb_plan_destroy(skb_plan *p);

/* ---------------------------------------------------------------- hot path (device pointers)
 * K1/K4  KinematicsMapBase.map  =  tinyfk.solve_fk         skmp/kinematics.py:43-63
 *        out_pts (N,F,T); out_jac (N,F,T,D) or NULL when with_jacobian is 0 */
SKB_API int skb_fk(const skb_plan *p, int32_t dtype, const void *q_dev, int64_t N, int32_t with_jacobian,
                   void *out_pts_dev, void *out_jac_dev, void *stream);
/* K1+K2 fused  CollFreeConst._evaluate                       skmp/constraint.py:287-374
 *        values = sdf(x) - radius - margin.
 *        SKB_MODE_ALL:     out_vals (N,S), out_jac (N,S,D)
 *        SKB_MODE_CLOSEST: out_vals (N,1), out_jac (N,1,D)   (argmin over spheres of sdf - radius)
 *        out_jac may be NULL (with_jacobian=False); out_vals may be NULL when only validity is
 *        wanted; out_valid (N bytes, nullable) = all(values > 0)  (is_valid, constraint.py:89-91) */
SKB_API int skb_collfree(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N,
                         double margin, int32_t mode, void *out_vals_dev, void *out_jac_dev,
                         uint8_t *out_valid_dev, void *stream);
/* CollFreeConst rows of an SQP trajectory stack, written the way the solver consumes them
 *        (skmp/solver/nlp_solver/trajectory_constraint.py:128-195: per waypoint evaluate_single, dense fill, conversion to
 *        CSC with a cached pattern :34-47): out_data (N, D, S) is the Jacobian block of every configuration TRANSPOSED --
 *        with the configurations of B trajectories laid out (B, n_wp, D) it IS the block-diagonal CSC `data` array of
 *        each trajectory (column (waypoint, dof) holds that waypoint's S rows).  out_vals (N,S) nullable.  D <= 32. */
SKB_API int skb_collfree_csc(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N,
                             double margin, void *out_vals_dev, void *out_data_dev, void *stream);
/* K3  PairWiseSelfCollFreeConst._evaluate = tinyfk.compute_inter_link_sqdists + threshold + row-min
 *        skmp/constraint.py:749-778.   values = |xi-xj|^2 - (ri+rj)^2
 *        SKB_MODE_ALL: (N,P), (N,P,D);  SKB_MODE_CLOSEST: (N,1), (N,1,D) */
SKB_API int skb_pairwise(const skb_plan *p, int32_t dtype, const void *q_dev, int64_t N, int32_t mode,
                         void *out_vals_dev, void *out_jac_dev, uint8_t *out_valid_dev, void *stream);

/* tinyfk.solve_com_fk + COMStabilityConst._evaluate           skmp/constraint.py:899-928
 *        com = sum_f w_f x_f / sum_f w_f over the plan's features (skb_plan_set_weights).
 *        SKB_COM_POINT:     out_vals (N,3) = com,          out_jac (N,3,D) (nullable)      [solve_com_fk]
 *        SKB_COM_STABILITY: out_vals (N,1) = -sdf(com),    out_jac (N,1,D) (nullable)      [the constraint; `s` required] */
SKB_API int skb_com(const skb_plan *p, const skb_sdf *s /* nullable for SKB_COM_POINT */, int32_t dtype, const void *q_dev,
                    int64_t N, int32_t what, void *out_vals_dev, void *out_jac_dev, void *stream);

/* One damped Gauss-Newton step for B seeds at once -- the linear-algebra half of the batched replacement of the
 * sequential SLSQP restarts in satisfy_by_optimization (skmp/satisfy.py:45-143):
 *   q_new[b] = clamp(q[b] + dq, lb, ub),  (J^T J + lam_b (I + diag J^T J)) dq = -J^T r
 * J (B,M,D), r (B,M), lam (B), q / q_new (B,D), lb / ub (D); all device pointers of `dtype`; 1 <= D <= 64. */
SKB_API int skb_lm_step(int32_t dtype, const void *J_dev, const void *r_dev, const void *lam_dev, const void *q_dev,
                        const void *lb_dev, const void *ub_dev, void *q_new_dev, int64_t B, int32_t M, int32_t D,
                        void *stream);

/* Batched edge discretisation for is_valid_motion_step (skmp/solver/motion_step_box.py:8-57), fp64 and sequential like
 * the reference so the sample lists are bit-identical:
 *   skb_edge_counts   counts[e] = len(interpolate_fractions(box, q1[e], q2[e], include_q1))
 *   skb_edge_samples  rows offsets[e] .. offsets[e+1] of `samples` = q1[e] + (q2[e] - q1[e]) * frac   (out_dtype rows)
 *   skb_segment_all   out[e] = AND of valid[offsets[e] .. offsets[e+1])   (an edge without samples is valid)
 * q1 / q2 (E,D) fp64, box (D) fp64, offsets (E+1) int64 = exclusive prefix sum of counts. */
SKB_API int skb_edge_counts(const double *q1_dev, const double *q2_dev, const double *box_dev, int64_t E, int32_t D,
                            int32_t include_q1, int64_t *counts_dev, void *stream);
SKB_API int skb_edge_samples(const double *q1_dev, const double *q2_dev, const double *box_dev, int64_t E, int32_t D,
                             int32_t include_q1, const int64_t *offsets_dev, int32_t out_dtype, void *samples_dev,
                             void *stream);
SKB_API int skb_segment_all(const uint8_t *valid_dev, const int64_t *offsets_dev, int64_t E, uint8_t *out_dev, void *stream);

/* Measurement aid: FMA-pipe peak of the current device in TFLOP/s, the roofline denominator of the
 * compute-bound modes (closest row, validity bytes).  kind 0: fp32 FFMA, 1: packed fma.rn.f32x2, 2: fp64 DFMA. */
SKB_API int skb_fma_peak(int32_t kind, int32_t iters, double *tflops_out);
/* Measurement aid: GB/s a write-only stream of `bytes` reaches on the current device -- the practical ceiling of the
 * write-dominated collision kernels (MEASURED_PEAKS.json's HBM figure is a read + write copy).  kind 0: 16-byte streaming
 * stores from registers; kind 1: the step kernel's store path, `block_bytes` blocks staged in shared memory and shipped by
 * cp.async.bulk (TMA 1-D) with the L2 evict-first policy.  No reference counterpart. */
SKB_API int skb_store_peak(int32_t kind, int64_t bytes, int32_t block_bytes, double *gbs_out);

/* ---------------------------------------------------------------- hot path (host pointers)
 * Same semantics with host buffers: what a numpy caller of the reference sees.  Chunked
 * H2D -> kernel -> D2H over internal streams; synchronous on return. */
SKB_API int skb_fk_host(const skb_plan *p, int32_t dtype, const void *q_host, int64_t N, int32_t with_jacobian,
                        void *out_pts_host, void *out_jac_host);
SKB_API int skb_collfree_host(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_host, int64_t N,
                              double margin, int32_t mode, void *out_vals_host, void *out_jac_host,
                              uint8_t *out_valid_host);
SKB_API int skb_pairwise_host(const skb_plan *p, int32_t dtype, const void *q_host, int64_t N, int32_t mode,
                              void *out_vals_host, void *out_jac_host, uint8_t *out_valid_host);
SKB_API int skb_com_host(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_host, int64_t N, int32_t what,
                         void *out_vals_host, void *out_jac_host);

/* number of kernel launches issued by this process through the library (bench.py's gpu_launches) */
SKB_API int64_t skb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SKMP_B200_H */

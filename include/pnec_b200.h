/*
 * pnec_b200.h — C-ABI of the B200-native PNEC frame-pair solver.
 *
 * This is the drop-in boundary for the reference's Ceres-backed refinement
 * (tum-vision/pnec).  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree).  Plain pointers and sizes
 * only: no C++/torch/Eigen types cross this boundary.
 *
 * Data layout (identical to what the reference already holds in memory):
 *   bearing vectors   double[n][3]   == std::vector<Eigen::Vector3d>::data()
 *   covariances       double[n][9]   == std::vector<Eigen::Matrix3d>::data()
 *                                       (column-major 3x3; the path only ever
 *                                       forms x^T S x, so only the symmetric
 *                                       part of S matters)
 *   pose              double[7]      qx qy qz qw tx ty tz
 *                                       == Sophus::SE3d memory order
 * A batch is the concatenation of B independent frame pairs; problem b owns
 * correspondences [offsets[b], offsets[b+1]).  `offsets == NULL` means a
 * uniform batch of `n_per_problem` correspondences each.
 */
#ifndef PNEC_B200_H_
#define PNEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNEC_B200_VERSION_MAJOR 0
#define PNEC_B200_VERSION_MINOR 2

/* Residual variants.
 *   NEC        include/optimization/nec_residual.h:47-69
 *   TARGET     include/optimization/pnec_residual.h:81-111  (PNEC::CeresSolver default)
 *   HOST       include/optimization/pnec_residual.h:50-79
 *   SYMMETRIC  include/optimization/pnec_residual.h:113-150 (pypnec.pyceres)
 */
typedef enum pnec_variant {
  PNEC_VARIANT_NEC = 0,
  PNEC_VARIANT_TARGET = 1,
  PNEC_VARIANT_HOST = 2,
  PNEC_VARIANT_SYMMETRIC = 3
} pnec_variant;

/* Where the batch pointers live. */
typedef enum pnec_memspace {
  PNEC_MEM_HOST = 0,  /* library stages H2D / D2H itself and synchronises   */
  PNEC_MEM_DEVICE = 1 /* all pointers are device pointers, call is async     */
} pnec_memspace;

/* Return codes of the entry points (0 == success). */
typedef enum pnec_error {
  PNEC_OK = 0,
  PNEC_ERR_INVALID_ARGUMENT = -1,
  PNEC_ERR_CUDA = -2,
  PNEC_ERR_NO_DEVICE = -3,
  PNEC_ERR_UNSUPPORTED = -4,
  PNEC_ERR_ALLOC = -5
} pnec_error;

/* Per-problem termination status — the information the reference keeps in
 * ceres::Solver::Summary (include/optimization/pnec_ceres.h:88) and never
 * exposes. */
typedef enum pnec_status {
  PNEC_STATUS_CONVERGED_FUNCTION = 0,  /* |dcost| <= function_tolerance*cost  */
  PNEC_STATUS_CONVERGED_PARAMETER = 1, /* |dx| <= parameter_tolerance*(|x|+tol) */
  PNEC_STATUS_CONVERGED_GRADIENT = 2,  /* max|g| <= gradient_tolerance        */
  PNEC_STATUS_CONVERGED_RADIUS = 3,    /* trust region radius <= min          */
  PNEC_STATUS_MAX_ITERATIONS = 4,      /* NO_CONVERGENCE in Ceres terms       */
  PNEC_STATUS_FAILURE = 5,             /* too many consecutive invalid steps  */
  PNEC_STATUS_NONFINITE = 6,           /* non-finite cost at the start point  */
  PNEC_STATUS_EMPTY = 7                /* zero correspondences: pose unchanged */
} pnec_status;

/* Solver options.  Field names and defaults are those of
 * ceres::Solver::Options as the reference uses it: always default-constructed
 * (src/optimization/pnec_ceres.cc:43-48, src/rel_pose_estimation/pnec.cc:355).
 * `regularization` is Options::regularization_
 * (include/rel_pose_estimation/pnec_config.h:50). */
typedef struct pnec_solver_opts {
  int32_t variant;                           /* pnec_variant                  */
  int32_t max_num_iterations;                /* 50                            */
  int32_t max_num_consecutive_invalid_steps; /* 5                             */
  int32_t jacobi_scaling;                    /* 1                             */
  double regularization;                     /* 1e-13                         */
  double function_tolerance;                 /* 1e-6                          */
  double gradient_tolerance;                 /* 1e-10                         */
  double parameter_tolerance;                /* 1e-8                          */
  double initial_trust_region_radius;        /* 1e4                           */
  double max_trust_region_radius;            /* 1e16                          */
  double min_trust_region_radius;            /* 1e-32                         */
  double min_relative_decrease;              /* 1e-3                          */
  double min_lm_diagonal;                    /* 1e-6                          */
  double max_lm_diagonal;                    /* 1e32                          */
} pnec_solver_opts;

/* Layout of the covariance arrays.  The path only ever forms x^T S x, so only sym(S) matters:
 * PACKED carries exactly that, (xx, (xy+yx)/2, (xz+zx)/2, yy, (yz+zy)/2, zz) per correspondence,
 * 48 bytes instead of 72 over the host link (96 instead of 120 per correspondence), and gives
 * bit-identical results (the kernels reduce a full matrix to the same six numbers on load). */
typedef enum pnec_cov_layout {
  PNEC_COV_FULL = 0,   /* double[n][9], column-major 3x3 == std::vector<Eigen::Matrix3d> */
  PNEC_COV_PACKED = 1  /* double[n][6], symmetric part                                   */
} pnec_cov_layout;

/* A batch of independent frame pairs. */
typedef struct pnec_batch {
  int64_t num_problems;      /* B                                             */
  int64_t n_per_problem;     /* uniform N when offsets == NULL                */
  const int64_t *offsets;    /* B+1 entries, HOST pointer always, or NULL     */
  int32_t memspace;          /* pnec_memspace of every pointer below          */
  int32_t cov_layout;        /* pnec_cov_layout of covs_target / covs_host    */
  const double *bvs_host;    /* [total][3]  f1, frame-1 ("host") bearings     */
  const double *bvs_target;  /* [total][3]  f2, frame-2 ("target") bearings   */
  const double *covs_target; /* [total][9]  TARGET, SYMMETRIC; HOST uses it too
                                (PNECCeres::Optimize(..., covs, reg, Host))  */
  const double *covs_host;   /* [total][9]  SYMMETRIC only, else NULL         */
  const double *poses;       /* [B][7]      start pose (solve) / eval pose    */
} pnec_batch;

/* Outputs of a batched solve.  Any pointer except `poses` may be NULL. */
typedef struct pnec_solve_out {
  double *poses;       /* [B][7] unit quaternion (x,y,z,w) + unit translation
                          == PNECCeres::Result(), src/optimization/pnec_ceres.cc:201-206 */
  int32_t *status;     /* [B] pnec_status                                     */
  int32_t *iterations; /* [B] LM iterations taken (Summary::iterations.size()-1) */
  double *cost;        /* [B] final cost 1/2 sum r^2 at the returned pose     */
  double *initial_cost;/* [B] cost at the start pose                          */
} pnec_solve_out;

/* Outputs of one fused evaluation (residual + Jacobian + J^T J reduction) at a
 * fixed pose per problem.  Tangent order (theta, phi, d1, d2, d3) follows the
 * parameter-block order of problem.AddResidualBlock(cf, nullptr, &theta_,
 * &phi_, q) + EigenQuaternionManifold, src/optimization/pnec_ceres.cc:99-106. */
typedef struct pnec_eval_out {
  double *cost;     /* [B]      1/2 sum r^2                                   */
  double *gradient; /* [B][5]   J^T r                                         */
  double *jtj;      /* [B][15]  upper triangle of J^T J, row-major packed     */
} pnec_eval_out;

typedef struct pnec_handle pnec_handle;

/* Library / error plumbing. */
int pnec_version(void);              /* major*1000 + minor                    */
const char *pnec_last_error(void);   /* thread-local message of the last failure */
const char *pnec_status_string(int32_t status);

/* ceres::Solver::Options() defaults + Options::regularization_ = 1e-13, variant TARGET. */
void pnec_solver_opts_default(pnec_solver_opts *opts);

/* A handle owns the device scratch and a stream-ordered staging area; it is
 * the analogue of one `PNECCeres optimizer;` local
 * (src/rel_pose_estimation/pnec.cc:355) and is re-entrant per handle. */
int pnec_create(int device, pnec_handle **out);
void pnec_destroy(pnec_handle *h);

/* Batched LM refinement, all iterations on device.
 * Replaces PNECCeres::Optimize (both overloads, src/optimization/pnec_ceres.cc:70-168)
 * and NECCeres::Optimize (src/optimization/nec_ceres.cc:73-101), i.e. the
 * bodies of PNEC::CeresSolver / CeresSolverFull / NECCeresSolver
 * (src/rel_pose_estimation/pnec.cc:350-411) and pypnec.pyceres / pyceresnec
 * (python/pypnec.cpp:50-82), for B frame pairs at once.
 * `cuda_stream` is a cudaStream_t (NULL = default stream).
 * HOST memspace: returns after results are in the caller's buffers.
 * DEVICE memspace: enqueues on the stream and returns. */
int pnec_solve_batch(pnec_handle *h, const pnec_batch *batch,
                     const pnec_solver_opts *opts, const pnec_solve_out *out,
                     void *cuda_stream);

/* One fused residual + analytic 5-DoF Jacobian + J^T J / J^T r / cost pass.
 * Replaces one ceres::Problem evaluation, i.e. N x
 * NumericDiffCostFunction<..., CENTRAL, 1,1,1,4>::Evaluate
 * (src/optimization/pnec_ceres.cc:92-101) plus the normal-equation assembly
 * inside ceres::Solve (:110). */
int pnec_eval_batch(pnec_handle *h, const pnec_batch *batch, int32_t variant,
                    double regularization, const pnec_eval_out *out,
                    void *cuda_stream);

/* Mean PNEC energy without regularisation at the given poses — the parity
 * metric pnec::common::CostFunction, src/common/common.cc:237-259. */
int pnec_cost_function_batch(pnec_handle *h, const pnec_batch *batch,
                             double *out_mean_energy /* [B] */,
                             void *cuda_stream);

/* Camera models of pnec::common::CameraModel (include/common/common.h:61). */
typedef enum pnec_camera_model {
  PNEC_CAMERA_OMNIDIRECTIONAL = 0,
  PNEC_CAMERA_PINHOLE = 1
} pnec_camera_model;

/* Unscented transform of n image-space covariances to 3x3 bearing-vector covariances:
 * the step that produces the `covs_*` inputs of the solve.
 * Replaces pnec::common::UnscentedTransform, both overloads
 * (src/common/common.cc:467-550; declared include/common/common.h:103-113):
 * 5 sigma points mu, mu +- C.col(i) with C the Cholesky factor of the 2x2 image-plane
 * block (rotated into the tangent plane for omnidirectional cameras), weights
 * kappa/(2+kappa) and 0.5/(2+kappa), projected with (K_inv p).normalized().
 *   mus   [n][3]  image-space points (mu)           covs [n][9] column-major
 *   K_inv [9]     column-major, HOST pointer        out  [n][9] column-major
 * `memspace` applies to mus / covs / out. */
int pnec_unscented_transform_batch(pnec_handle *h, int64_t n, int32_t memspace,
                                   const double *mus, const double *covs,
                                   const double *K_inv, double kappa, int32_t camera_model,
                                   double *out_covs, void *cuda_stream);

/* Keypoints to solver inputs: bearing vector + 3x3 bearing covariance of n keypoints from
 * their pixel position and 2x2 image covariance, i.e. KeyPoint::Unproject
 * (src/frames/keypoints.cc:49-62): bearing = Unproject(point, K_inv)
 * (src/common/common.cc:460-465) and covariance = UnscentedTransform((x, y, 1),
 * [[cov2, 0], [0, 0]], K_inv, 1.0, Pinhole).  The solver's inputs can then be produced on the
 * device from 6 doubles per keypoint instead of being shipped as 12.
 *   points [n][2]   covs2 [n][4] column-major 2x2 (Eigen::Matrix2d)   K_inv [9] HOST pointer
 *   out_bvs [n][3]  out_covs [n][9] column-major */
int pnec_keypoints_unproject_batch(pnec_handle *h, int64_t n, int32_t memspace,
                                   const double *points, const double *covs2,
                                   const double *K_inv, double *out_bvs, double *out_covs,
                                   void *cuda_stream);

/* Translation given rotation, probabilistic: the SCF stage of PNEC::WeightedEigensolver
 * (src/rel_pose_estimation/pnec.cc:317-343) for B frame pairs.  Per pair, with R the rotation
 * and t0 the translation of `batch->poses`:
 *   A_i = n_i n_i^T, n_i = f1 x R f2;   B_i = [f1]x R S_i R^T [f1]x^T + regularization * I
 *   start = first strict minimiser of sum_i t^T A_i t / t^T B_i t over {t0} followed by
 *           fibonacci_sphere(fibonacci_samples)            (src/optimization/scf.cc:43-72)
 *   then `scf_steps` iterations t <- eigenvector of the smallest eigenvalue of
 *           E(t) = sum_i A_i / (t^T B_i t)                 (scf.cc:109-147, including the
 *           `frac.resize(n)` + push_back quirk that makes frac[i] == 0)
 * The reference calls it with 500 samples and 10 steps.  The sign of the result is that of the
 * eigenvector routine (arbitrary, as with Eigen): compare modulo sign.  Pairs up to ~3000
 * correspondences are processed out of shared memory; larger ones keep their per-correspondence
 * terms in a device scratch array (slower, same arithmetic).
 *   out_translations [B][3]      out_cost [B] objective at the result, or NULL */
int pnec_scf_translation_batch(pnec_handle *h, const pnec_batch *batch, double regularization,
                               int32_t fibonacci_samples, int32_t scf_steps,
                               double *out_translations, double *out_cost, void *cuda_stream);

/* Translation given rotation, NEC: TranslationFromM(ComposeM(bvs_1, bvs_2, R))
 * (src/common/common.cc:127-181; the translation of PNEC::Eigensolver, pnec.cc:270-278):
 * M = sum_{i >= 1} n_i n_i^T (ComposeM skips the first correspondence, common.cc:131) and the
 * unit eigenvector of its smallest eigenvalue (sign arbitrary).
 *   out_translations [B][3]      out_M [B][6] (xx xy xz yy yz zz) or NULL */
int pnec_nec_translation_batch(pnec_handle *h, const pnec_batch *batch, double *out_translations,
                               double *out_M, void *cuda_stream);

/* Rotation by NEC eigenvalue minimisation for B frame pairs:
 *   rotation_t opengv::relative_pose::eigensolver(CentralRelativeAdapter(bvs1, bvs2, R12))
 * as the reference calls it at src/rel_pose_estimation/pnec.cc:274 (PNEC::Eigensolver, no RANSAC)
 * and pnec.cc:313 (PNEC::WeightedEigensolver).  opengv is a third-party dependency outside the
 * reference tree; what is computed (restated from its publication, Kneip & Lynen ICCV 2013, and its
 * public sources): the moments sum_i w_i f1_u f1_v f2 f2^T, then Levenberg-Marquardt (Eigen's MINPACK
 * lmdif port: forward differences, ftol 5e-5, xtol 10 eps, maxfev 100) on the gradient of the smallest
 * eigenvalue of M(c) = sum_i n_i n_i^T, n_i = f1_i x R'(c) f2_i, over the Cayley parameters c,
 * started at the rotation of `batch->poses`.
 *   weight_poses  NULL: unweighted (PNEC::Eigensolver).  Otherwise [B][7] (same memspace): f2_i is
 *                 scaled by sqrt(Weight_i * 1e-8), Weight = 1 / (t^T [f1]x R S_i R^T [f1]x^T t + reg)
 *                 at these poses (pnec::common::Weight, src/common/common.cc:183-208 with
 *                 host_frame = false; pnec.cc:294-306); needs batch->covs_target.
 *   out_poses     [B][7]: unit quaternion of the result; the translation of batch->poses is passed
 *                 through unchanged
 *   out_lm_info   [B] MINPACK termination code (1..8), or NULL
 *   out_smallest_ev [B] smallest eigenvalue of opengv's (unnormalised) M at the result, or NULL */
int pnec_eigensolver_batch(pnec_handle *h, const pnec_batch *batch, const double *weight_poses,
                           double regularization, double *out_poses, int32_t *out_lm_info,
                           double *out_smallest_ev, void *cuda_stream);

/* pnec::rel_pose_estimation::Options as PNEC::Solve reads it
 * (include/rel_pose_estimation/pnec_config.h:46-65), plus the literals of the call sites. */
typedef struct pnec_frame_opts {
  int32_t use_nec;             /* use_nec_              false                                   */
  int32_t use_ceres;           /* use_ceres_            true                                    */
  int32_t weighted_iterations; /* weighted_iterations_  10                                      */
  int32_t use_ransac;          /* use_ransac_           true  (pnec_config.h:58)                */
  int32_t fibonacci_samples;   /* 500, the literal at pnec.cc:331                               */
  int32_t scf_steps;           /* 10,  the literal at pnec.cc:342                               */
  pnec_solver_opts ceres;      /* ceres_options_ (+ regularization_); `variant` is ignored: NEC
                                  when use_nec, TARGET otherwise                                */
  int32_t max_ransac_iterations; /* max_ransac_iterations_  5000 (pnec_config.h:59)             */
  int32_t ransac_sample_size;    /* ransac_sample_size_     10   (pnec_config.h:60), at most 32 */
  double ransac_threshold;       /* 1e-6, the literal at pnec.cc:250                            */
  double ransac_probability;     /* 0.99, opengv::sac::Ransac's default probability_            */
  double ransac_max_variation;   /* 0.1, the start perturbation of
                                    EigensolverSacProblem::computeModelCoefficients             */
  uint64_t ransac_seed;          /* key of the counter-based random stream (opengv: time-seeded
                                    mt19937 + rand(), not reproducible)                         */
  int64_t ransac_pair_index_base;/* 0; pair b draws from the stream of pair (base + b): a shard of
                                    a larger batch passes its first pair's index and reproduces
                                    what the whole batch would compute                          */
} pnec_frame_opts;

/* Options() defaults (use_ransac = 1). */
void pnec_frame_opts_default(pnec_frame_opts *opts);

typedef struct pnec_frame_out {
  double *poses;       /* [B][7] result of PNEC::Solve                                          */
  double *es_poses;    /* [B][7] result of PNEC::Eigensolver (ES_solution, pnec.cc:86), or NULL */
  int32_t *status;     /* [B] pnec_status of the refinement, or NULL (untouched if !use_ceres)  */
  int32_t *iterations; /* [B] or NULL                                                           */
  double *cost;        /* [B] or NULL                                                           */
  /* RANSAC (`std::vector<int> &inliers` of PNEC::Solve, pnec.cc:81): any may be NULL.  Without
   * RANSAC num_inliers is 0 (the reference clears the vector, pnec.cc:277).                    */
  int32_t *num_inliers;       /* [B]                                                            */
  int32_t *inlier_index;      /* [total] indices within the pair, ascending; pair b's list starts
                                 at its first correspondence's position and has num_inliers[b]
                                 entries                                                        */
  int32_t *ransac_iterations; /* [B] opengv's ransac.iterations_                                */
  /* Stage timings of the timed Solve overloads (pnec.cc:145-205; FrameTiming::nec_es_, it_es_,
   * ceres_, include/common/timing.h:52-55) from CUDA events, milliseconds, HOST pointer [3], or
   * NULL.  Asking for them makes the call synchronous and runs the batch as one chunk.         */
  float *stage_ms;
} pnec_frame_out;

/* opengv::sac::Ransac<EigensolverSacProblem>::computeModel + selectWithinDistance for B frame
 * pairs: the RANSAC stage of PNEC::Eigensolver alone (src/rel_pose_estimation/pnec.cc:239-251),
 * exposed for testing.  Uses use_ransac's settings of `opts` (max_ransac_iterations,
 * ransac_sample_size, ransac_threshold, ransac_probability, ransac_max_variation, ransac_seed).
 * Every hypothesis is a function of (seed, pair index, iteration) alone -- a fresh partial
 * Fisher-Yates sample and a start at the rotation of batch->poses moved by U(-1, 1) *
 * ransac_max_variation per Cayley parameter -- so rounds of hypotheses are evaluated in parallel
 * and opengv's bookkeeping (first strict maximum of the inlier count, k = log(1 - p) /
 * log(1 - w^s)) is replayed over them in order.  opengv's own loop additionally carries a
 * shuffled index array and the previous model's rotation from one iteration to the next
 * (oracle/pnec_oracle_frame.c, `sequential`): the same distribution of samples, not the same
 * stream; parity with the reference is statistical by construction (it seeds from time(0)).
 *   out_models        [B][7] winning hypothesis: unit quaternion + unit translation (signed by
 *                     the optical flow of the sample's first correspondence); the start pose when a
 *                     pair has fewer correspondences than the sample size
 *   out_num_inliers   [B]      out_iterations [B] or NULL
 *   out_inlier_index  [total] (layout as in pnec_frame_out) or NULL
 *   pair_index_base   pair b draws from the stream of pair (pair_index_base + b)               */
int pnec_ransac_batch(pnec_handle *h, const pnec_batch *batch, const pnec_frame_opts *opts,
                      int64_t pair_index_base, double *out_models, int32_t *out_num_inliers,
                      int32_t *out_iterations, int32_t *out_inlier_index, void *cuda_stream);

/* The whole frame-to-frame solve for B frame pairs, every stage on the device:
 * Sophus::SE3d PNEC::Solve(bvs1, bvs2, projected_covs, initial_pose, inliers)
 * (src/rel_pose_estimation/pnec.cc:77-124):
 *   0. use_ransac: pnec_ransac_batch, then InlierExtraction (pnec.cc:210-229): every later stage
 *      sees the inliers only; step 1 then is optimizeModelCoefficients (pnec.cc:253-256): the
 *      eigensolver over the inliers started at the winning hypothesis
 *   1. PNEC::Eigensolver (pnec.cc:231-281): pnec_eigensolver_batch from the rotation of
 *      batch->poses, translation = TranslationFromM(ComposeM(..)) (pnec_nec_translation_batch)
 *   2. use_nec: NECCeresSolver from 1 (or 1 itself if !use_ceres)
 *   3. else weighted_iterations > 1: PNEC::WeightedEigensolver (pnec.cc:283-348):
 *      (weighted_iterations - 1) x { weighted eigensolver started at the previous rotation, weights
 *      from the pose of step 1; SCF translation started at the previous translation };
 *      == 1: the pose of step 1;  == 0: batch->poses
 *   4. use_ceres: CeresSolver (TARGET residual) from 3
 * batch->covs_target may be NULL when use_nec. */
int pnec_frame_solve_batch(pnec_handle *h, const pnec_batch *batch, const pnec_frame_opts *opts,
                           const pnec_frame_out *out, void *cuda_stream);

/* ---- Solver inputs from keypoints: the frame-level boundary.
 *
 * In the reference the solver's inputs are assembled per frame pair by Frame2Frame::GetFeatures
 * (src/rel_pose_estimation/frame2frame.cc:359-392) from the matched keypoints of the two frames,
 * whose bearing vector and 3x3 covariance were derived at construction from the pixel position
 * and the 2x2 image covariance (KeyPoint::Unproject, src/frames/keypoints.cc:49-62).  Here that
 * derivation runs on the device for exactly the matched keypoints, so a caller hands over what a
 * KeyPoint is constructed from: 16 B (host keypoint) + 48 B (target keypoint) per correspondence
 * instead of 120 B -- the host link is what bounds an end-to-end call.
 *
 * Two keypoint tables (host frame(s) / target frame(s)); correspondence i of the batch uses row
 * host_index[i] / target_index[i] (match.queryIdx / match.trainIdx resolved to table rows), or row
 * i when the index array is NULL.  Covariances: 2x2 column-major (Eigen::Matrix2d, [K][4]) or,
 * with packed_covs, their symmetric part (xx, xy, yy), [K][3].  The target table's covariances
 * give covs_target (noise frame Target, the reference's default); host_covs2 is only needed
 * for the SYMMETRIC variant. */
typedef struct pnec_keypoint_batch {
  int64_t num_problems;        /* B frame pairs                                           */
  int64_t n_per_problem;       /* uniform N when offsets == NULL                          */
  const int64_t *offsets;      /* [B+1] HOST pointer, or NULL                             */
  int32_t memspace;            /* pnec_memspace of every pointer below except K_inv       */
  int32_t packed_covs;         /* 0: [K][4] column-major 2x2;  1: [K][3] (xx, xy, yy)     */
  int64_t num_host_keypoints;  /* rows of host_points (/ host_covs2)                      */
  int64_t num_target_keypoints;/* rows of target_points / target_covs2                    */
  const double *host_points;   /* [Kh][2] KeyPoint::point_                                */
  const double *target_points; /* [Kt][2]                                                 */
  const double *host_covs2;    /* [Kh][4|3] KeyPoint::img_covariance_, or NULL            */
  const double *target_covs2;  /* [Kt][4|3], or NULL for NEC-only use                     */
  const int32_t *host_index;   /* [total] or NULL                                         */
  const int32_t *target_index; /* [total] or NULL                                         */
  const double *K_inv;         /* [9] column-major, HOST pointer                          */
  const double *poses;         /* [B][7] start poses                                      */
} pnec_keypoint_batch;

/* Builds the solver's inputs on the device (one pass: unprojection + unscented transform of the
 * matched keypoints) and returns them as a DEVICE-memspace pnec_batch in *out_batch, usable with
 * every entry point above.  The arrays belong to the handle and stay valid until the next call
 * on it that stages a HOST batch or keypoints.  Enqueued on `cuda_stream`. */
int pnec_keypoints_to_batch(pnec_handle *h, const pnec_keypoint_batch *kb, pnec_batch *out_batch,
                            void *cuda_stream);

/* pnec_solve_batch / pnec_frame_solve_batch fed from keypoints.  `out` pointers live in
 * kb->memspace.  HOST: the keypoint tables are cut into chunks whose copy, assembly, solve and
 * result copy overlap; returns with the results in the caller's buffers. */
int pnec_solve_from_keypoints_batch(pnec_handle *h, const pnec_keypoint_batch *kb,
                                    const pnec_solver_opts *opts, const pnec_solve_out *out,
                                    void *cuda_stream);
int pnec_frame_solve_from_keypoints_batch(pnec_handle *h, const pnec_keypoint_batch *kb,
                                          const pnec_frame_opts *opts, const pnec_frame_out *out,
                                          void *cuda_stream);

/* Number of kernels this handle has launched so far (bench bookkeeping). */
int64_t pnec_launch_count(const pnec_handle *h);

#ifdef __cplusplus
}
#endif

#endif /* PNEC_B200_H_ */

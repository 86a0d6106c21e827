/*
 * clc_oracle.h -- CPU ORACLE for the camera<->laser extrinsic solve.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithm of the reference hot path
 *   /root/reference/src/LaseCamCalCeres.cpp            (cost model, problem assembly, closed form, analysis tail)
 *   /root/reference/src/pose_local_parameterization.cpp (SE(3) "plus")
 *   /root/reference/src/utilities.cpp:267-272           (pi_from_ppp)
 *   /root/reference/main/calibr_simulation.cpp:10-108   (synthetic generator)
 * plus the published trust-region / Levenberg-Marquardt semantics of Ceres Solver (<= 2.1; un-vendored and
 * un-pinned by the reference, see CMakeLists.txt:37), which the reference drives through ceres::Solve().
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path, and neither Ceres
 * nor Eigen exists in this image, so the reference itself cannot be compiled or run here (DESIGN.md section
 * "Oracle").  The only reference-pinned answer is semantic: noise-free simulated data is solved exactly at the
 * printed ground truth (main/calibr_simulation.cpp:15-25).  This oracle is cross-checked against an independent
 * numpy twin (oracle/oracle_np.py), sympy derivatives and scipy.optimize (tests/).
 *
 * Nothing under oracle/ may be imported, linked or executed by the product (camlasercalibratool_b200/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * and as the timed CPU baseline.
 *
 * Conventions
 *   pose7      = (tx,ty,tz,qx,qy,qz,qw)   the Ceres parameter block, LaseCamCalCeres.cpp:219
 *   frame_pose = (qx,qy,qz,qw,tx,ty,tz)   Oberserve::tagPose_Qca (Eigen coeff order) then tagPose_tca,
 *                                         include/LaseCamCalCeres.h:20-21
 *   points     = AoS xyz (Eigen::Vector3d), CSR offsets[n_frames+1] delimit the frames
 *   edge_points= per frame 6 doubles: obi.points.front() then obi.points.back(), or NULL  (LaseCamCalCeres.cpp:278-279)
 */
#ifndef CLC_ORACLE_H
#define CLC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int64_t n_frames;
  const double* frame_pose;   /* [n_frames*7] */
  const int64_t* offsets;     /* [n_frames+1] */
  const double* points;       /* [offsets[n_frames]*3]  the calibration point set chosen by use_linefitting_data */
  const double* edge_points;  /* [n_frames*6] or NULL: enables the boundary residuals of LaseCamCalCeres.cpp:258-294 */
  int use_loss;               /* 1 = CauchyLoss(0.05*scale) as in the reference (#define LOSSFUNCTION, :212) */
  double cauchy_a;            /* 0.05 (:249) */
} oracle_problem;

/* Ceres Solver::Options fields that matter on this path; oracle_default_options() fills the Ceres defaults
 * overridden by LaseCamCalCeres.cpp:302-304 (DENSE_QR, max_num_iterations = 100). */
typedef struct {
  int max_num_iterations;
  double initial_trust_region_radius;
  double max_trust_region_radius;
  double min_trust_region_radius;
  double min_relative_decrease;
  double min_lm_diagonal;
  double max_lm_diagonal;
  double function_tolerance;
  double gradient_tolerance;
  double parameter_tolerance;
  int max_num_consecutive_invalid_steps;
  int jacobi_scaling;
  int linear_solver;  /* 0 = DENSE_QR on the materialised Jacobian (what the reference does),
                         1 = normal equations + Cholesky (streaming; never materialises J) */
  int num_threads;    /* OpenMP threads for the residual sweeps; 1 = what the reference uses */
} oracle_options;

enum {
  ORACLE_TERM_CONVERGENCE_FUNCTION = 1,
  ORACLE_TERM_CONVERGENCE_PARAMETER = 2,
  ORACLE_TERM_CONVERGENCE_GRADIENT = 3,
  ORACLE_TERM_CONVERGENCE_MIN_RADIUS = 4,
  ORACLE_TERM_NO_CONVERGENCE = 5, /* max_num_iterations */
  ORACLE_TERM_FAILURE = 6         /* too many invalid steps / evaluation failure */
};

/* One row of Ceres' IterationSummary (the fields FullReport prints). */
typedef struct {
  int iteration;
  int step_is_valid;
  int step_is_successful;
  double cost;
  double cost_change;
  double gradient_max_norm;
  double step_norm;
  double relative_decrease;
  double trust_region_radius;
} oracle_iteration;

typedef struct {
  int termination;
  int num_iterations; /* rows written to trace (iteration 0 included) */
  int num_successful_steps;
  int num_unsuccessful_steps;
  int num_residual_evaluations; /* sweeps over all residuals, with or without the Jacobian */
  int num_jacobian_evaluations;
  double initial_cost;
  double final_cost;
} oracle_summary;

void oracle_default_options(oracle_options* o);

/* Eigen restatements ---------------------------------------------------------------------------------- */
void oracle_quat_to_rot(const double q_xyzw[4], double R[9]);         /* QuaternionBase::toRotationMatrix */
void oracle_rot_to_quat(const double R[9], double q_xyzw[4]);         /* Quaternion(Matrix3) ctor */
void oracle_T_to_pose7(const double T_rowmajor[16], double pose7[7]); /* LaseCamCalCeres.cpp:215-219 */
void oracle_pose7_to_T(const double pose7[7], double T_rowmajor[16]); /* :311-314 (bottom row 0 0 0 1) */
void oracle_pose_plus(const double x[7], const double delta[6], double x_plus[7]); /* pose_local_parameterization.cpp:15-32 */

/* a2 / a7: planes --------------------------------------------------------------------------------------- */
void oracle_frame_plane(const double frame_pose[7], double plane[4]);              /* :227-231 */
void oracle_edge_planes(const double frame_pose[7], double pi1[4], double pi2[4]); /* :262-276 */

/* a3: one PointInPlaneFactor::Evaluate (:43-66).  jac7 may be NULL. */
void oracle_factor_evaluate(const double plane[4], const double pt[3], double scale, const double pose7[7],
                            double* residual, double* jac7);

/* Number of residuals of the problem (points + 2 per non-empty frame when edges are on). */
int64_t oracle_num_residuals(const oracle_problem* p);

/* Ceres-shaped evaluation: cost = 1/2 sum rho; residuals (loss-corrected) [R]; jacobian (loss-corrected,
 * local 6 columns, row-major [R*6]); gradient = J^T r [6].  Any output may be NULL. */
int oracle_evaluate(const oracle_problem* p, const double pose7[7], double* cost, double* residuals,
                    double* jacobian, double* gradient, int num_threads);

/* Streaming evaluation: accumulates H = J^T J (row-major 6x6), g = J^T r and cost in one pass without
 * materialising anything.  H and g may be NULL (cost only). */
int oracle_evaluate_normal(const oracle_problem* p, const double pose7[7], double* cost, double* H36, double* g6,
                           int num_threads);

/* a9: the Ceres trust-region LM solve.  trace may be NULL; trace_cap rows are available. */
int oracle_solve(const oracle_problem* p, double pose7[7], const oracle_options* opt, oracle_summary* summary,
                 oracle_iteration* trace, int trace_cap);

/* a11: analysis tail (:318-381): H = sum J^T J, b = -sum J^T r, chi = sum r^2 with the scale kept, no loss,
 * no edge residuals; singular values of H (descending). */
int oracle_information(const oracle_problem* p, const double pose7[7], double* H36, double* b6, double* chi,
                       double* singular_values6);

/* a12: closed-form initialisation (:112-203).  Uses x,y of the given point set. Tlc row-major 4x4.
 * AtA (row-major 9x9) and Atb (9) are returned for parity checks when non-NULL. */
int oracle_closed_form(const oracle_problem* p, double Tlc16[16], int* unobservable, double* AtA81, double* Atb9);

/* LineFittingCeres (LaseCamCalCeres.cpp:385-433): robust fit of m0 x + m1 y + 1 = 0 to the x,y of `points` (AoS xyz,
 * z ignored), CauchyLoss(0.05), DENSE_QR, max_num_iterations = 10, every other option a Ceres default; `line` is the
 * start value on entry (the reference's caller passes it uninitialised, calibr_offline.cpp:123) and the result on exit. */
int oracle_line_fit(const double* points, int64_t n, double line[2], int max_num_iterations, oracle_summary* summary,
                    oracle_iteration* trace, int trace_cap);

/* TranScanToPoints (src/utilities.cpp:181-215): float ranges -> xyz points, invalid beams -> (1000,1000,0). */
void oracle_scan_to_points(const float* ranges, int64_t n, double angle_min, double angle_increment, double range_min,
                           double* points);
/* AutoGetLinePts (src/selectScanPoints.cpp:17-190) without the OpenCV drawing: the longest continuous segment in the
 * +-80 degree front sector.  Returns 1 and the inclusive index range [start, end] of the chosen segment, else 0. */
int oracle_auto_get_line_pts(const double* points, int64_t n, int64_t* start, int64_t* end);

/* Small dense helpers exposed for the tests: singular values (descending) of a symmetric n x n matrix (n <= 9). */
void oracle_sym_singular_values(const double* A, int n, double* sv);

/* Synthetic generator (calibr_simulation.cpp:10-108 with a counter-based RNG) --------------------------- */
typedef struct {
  int64_t n_frames;
  int64_t beams;       /* M: beams per scan (reference: 180) */
  uint64_t seed;
  double sigma;        /* range noise along the ray, metres (reference: 0) */
  int exact_m;         /* 0 = faithful ragged frames; 1 = every frame has exactly `beams` points */
  int with_edges;      /* 1 = also emit edge_points consistent with the board-edge planes at ground truth */
} oracle_gen_desc;

/* Ground truth of the generator: T_lc (calibr_simulation.cpp:15-20) and its inverse T_cl as pose7. */
void oracle_gen_ground_truth(double Tlc16[16], double Tcl_pose7[7]);

/* Pass 1: returns the total point count and fills offsets[n_frames+1] and frame_pose[n_frames*7].
 * Pass 2 (oracle_gen_points) fills points[P*3] (and edge_points[n_frames*6] when with_edges). */
int64_t oracle_gen_frames(const oracle_gen_desc* g, double* frame_pose, int64_t* offsets);
int oracle_gen_points(const oracle_gen_desc* g, const double* frame_pose, const int64_t* offsets, double* points,
                      double* edge_points);

/* Philox4x32-10 block, exposed so that tests can pin the RNG against the CUDA twin. */
void oracle_philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* CLC_ORACLE_H */

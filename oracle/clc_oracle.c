/*
 * clc_oracle.c -- CPU ORACLE (test infrastructure, see clc_oracle.h).  PARITY UNPINNED (no reference tests /
 * golden vectors exist and Ceres+Eigen are absent from this image).
 *
 * Every function cites the reference lines it restates.  "Ceres:" citations name the public Ceres Solver
 * source file (<= 2.1) whose published algorithm is restated; Ceres is an un-vendored, un-pinned dependency
 * of the reference (CMakeLists.txt:37) and is not available in this image.
 */
#include "clc_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------------ */
/* Eigen restatements                                                                                     */
/* ------------------------------------------------------------------------------------------------------ */

/* Eigen::QuaternionBase::toRotationMatrix (Eigen/src/Geometry/Quaternion.h); no normalisation. Row-major R. */
void oracle_quat_to_rot(const double q[4], double R[9]) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.0 - (txx + tyy);
}

/* Eigen::Quaternion(const Matrix3&) -- Shoemake's method as in Eigen's quaternionbase_assign_impl<Other,3,3>.
 * Used by the reference at LaseCamCalCeres.cpp:215 and calibr_simulation.cpp:60. */
void oracle_rot_to_quat(const double R[9], double q[4]) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 3 + i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 3 + i] - R[j * 3 + j] - R[k * 3 + k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

/* LaseCamCalCeres.cpp:215-219: q from the rotation block, pose = (t, qx,qy,qz,qw). */
void oracle_T_to_pose7(const double T[16], double pose7[7]) {
  double R[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]};
  double q[4];
  oracle_rot_to_quat(R, q);
  pose7[0] = T[3]; pose7[1] = T[7]; pose7[2] = T[11];
  pose7[3] = q[0]; pose7[4] = q[1]; pose7[5] = q[2]; pose7[6] = q[3];
}

/* LaseCamCalCeres.cpp:311-314. */
void oracle_pose7_to_T(const double pose7[7], double T[16]) {
  double R[9];
  oracle_quat_to_rot(pose7 + 3, R);
  T[0] = R[0]; T[1] = R[1]; T[2] = R[2];  T[3] = pose7[0];
  T[4] = R[3]; T[5] = R[4]; T[6] = R[5];  T[7] = pose7[1];
  T[8] = R[6]; T[9] = R[7]; T[10] = R[8]; T[11] = pose7[2];
  T[12] = 0.0; T[13] = 0.0; T[14] = 0.0;  T[15] = 1.0;
}

/* pose_local_parameterization.cpp:3-32: p = p + dp; q = (q * [w=1, xyz=dtheta/2]).normalized(). */
void oracle_pose_plus(const double x[7], const double d[6], double xp[7]) {
  xp[0] = x[0] + d[0]; xp[1] = x[1] + d[1]; xp[2] = x[2] + d[2];
  const double ax = x[3], ay = x[4], az = x[5], aw = x[6];
  const double bx = d[3] / 2.0, by = d[4] / 2.0, bz = d[5] / 2.0, bw = 1.0;
  /* Eigen quaternion product a*b */
  double w = aw * bw - ax * bx - ay * by - az * bz;
  double xx = aw * bx + ax * bw + ay * bz - az * by;
  double yy = aw * by + ay * bw + az * bx - ax * bz;
  double zz = aw * bz + az * bw + ax * by - ay * bx;
  const double nrm = sqrt(xx * xx + yy * yy + zz * zz + w * w);
  xp[3] = xx / nrm; xp[4] = yy / nrm; xp[5] = zz / nrm; xp[6] = w / nrm;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Planes                                                                                                 */
/* ------------------------------------------------------------------------------------------------------ */

/* LaseCamCalCeres.cpp:227-231: planar_cam = (Tctag^-1)^T (0,0,1,0).  For the affine Tctag = [A t; 0 1] the
 * inverse is [A^-1, -A^-1 t], so the result is (row 2 of A^-1, -(row 2 of A^-1) . t).  A = R(Qca) is not
 * forced to be orthonormal (a non-unit quaternion gives a scaled A), hence the general 3x3 inverse. */
void oracle_frame_plane(const double fp[7], double plane[4]) {
  double A[9];
  oracle_quat_to_rot(fp, A);
  /* cofactors giving row 2 of A^-1:  inv(2,j) = C(j,2)/det */
  const double c0 = A[3] * A[7] - A[4] * A[6];   /* C(0,2) */
  const double c1 = -(A[0] * A[7] - A[1] * A[6]); /* C(1,2) */
  const double c2 = A[0] * A[4] - A[1] * A[3];   /* C(2,2) */
  const double det = A[2] * c0 + A[5] * c1 + A[8] * c2;
  const double n0 = c0 / det, n1 = c1 / det, n2 = c2 / det;
  plane[0] = n0; plane[1] = n1; plane[2] = n2;
  plane[3] = -(n0 * fp[4] + n1 * fp[5] + n2 * fp[6]);
}

/* utilities.cpp:267-272 pi_from_ppp. */
static void pi_from_ppp(const double x1[3], const double x2[3], const double x3[3], double pi[4]) {
  const double a[3] = {x1[0] - x3[0], x1[1] - x3[1], x1[2] - x3[2]};
  const double b[3] = {x2[0] - x3[0], x2[1] - x3[1], x2[2] - x3[2]};
  pi[0] = a[1] * b[2] - a[2] * b[1];
  pi[1] = a[2] * b[0] - a[0] * b[2];
  pi[2] = a[0] * b[1] - a[1] * b[0];
  const double c[3] = {x1[1] * x2[2] - x1[2] * x2[1], x1[2] * x2[0] - x1[0] * x2[2], x1[0] * x2[1] - x1[1] * x2[0]};
  pi[3] = -(x3[0] * c[0] + x3[1] * c[1] + x3[2] * c[2]);
}

/* LaseCamCalCeres.cpp:262-276: board corners in the tag frame -> camera frame -> the two planes through the
 * optical centre and a board edge (normals NOT normalised, d = 0). */
void oracle_edge_planes(const double fp[7], double pi1[4], double pi2[4]) {
  const double orig = 0.0265 + 0.0165;
  double pm[3][3] = {{0.0, 0.0, 0.0}, {0.5, 0.0, 0.0}, {0.0, 0.5, 0.0}};
  double R[9], pc[3][3];
  const double zero[3] = {0.0, 0.0, 0.0};
  oracle_quat_to_rot(fp, R);
  for (int k = 0; k < 3; ++k) {
    pm[k][0] -= orig; pm[k][1] -= orig; pm[k][2] -= 0.0;
    for (int r = 0; r < 3; ++r)
      pc[k][r] = (R[r * 3 + 0] * pm[k][0] + R[r * 3 + 1] * pm[k][1] + R[r * 3 + 2] * pm[k][2]) + fp[4 + r];
  }
  pi_from_ppp(pc[0], pc[1], zero, pi1);
  pi_from_ppp(pc[0], pc[2], zero, pi2);
}

/* ------------------------------------------------------------------------------------------------------ */
/* Cost model                                                                                             */
/* ------------------------------------------------------------------------------------------------------ */

/* PointInPlaneFactor::Evaluate, LaseCamCalCeres.cpp:43-66 (R is recomputed from q as the reference does). */
void oracle_factor_evaluate(const double plane[4], const double pt[3], double scale, const double pose7[7],
                            double* residual, double* jac7) {
  double R[9];
  oracle_quat_to_rot(pose7 + 3, R);
  const double pc0 = (R[0] * pt[0] + R[1] * pt[1] + R[2] * pt[2]) + pose7[0];
  const double pc1 = (R[3] * pt[0] + R[4] * pt[1] + R[5] * pt[2]) + pose7[1];
  const double pc2 = (R[6] * pt[0] + R[7] * pt[1] + R[8] * pt[2]) + pose7[2];
  *residual = scale * ((plane[0] * pc0 + plane[1] * pc1 + plane[2] * pc2) + plane[3]);
  if (jac7) {
    /* -R * skew(p):  skew(p) = [0 -pz py; pz 0 -px; -py px 0]  (:35-42) */
    double M[9];
    const double S[9] = {0.0, -pt[2], pt[1], pt[2], 0.0, -pt[0], -pt[1], pt[0], 0.0};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        M[r * 3 + c] = -(R[r * 3 + 0] * S[0 * 3 + c] + R[r * 3 + 1] * S[1 * 3 + c] + R[r * 3 + 2] * S[2 * 3 + c]);
    for (int c = 0; c < 3; ++c) {
      jac7[c] = scale * plane[c];
      jac7[3 + c] = scale * (plane[0] * M[0 * 3 + c] + plane[1] * M[1 * 3 + c] + plane[2] * M[2 * 3 + c]);
    }
    jac7[6] = 0.0;
  }
}

/* Ceres: loss_function.cc CauchyLoss::Evaluate with b = a*a, c = 1/b. */
static inline void cauchy_loss(double a, double s, double rho[3]) {
  const double b = a * a, c = 1.0 / b;
  const double sum = 1.0 + s * c;
  const double inv = 1.0 / sum;
  rho[0] = b * log(sum);
  rho[1] = inv > DBL_MIN ? inv : DBL_MIN;
  rho[2] = -c * (inv * inv);
}

/* Ceres: residual_block.cc ResidualBlock::Evaluate + corrector.cc.  For Cauchy rho'' < 0 always, so the
 * Corrector takes its simple branch: residual and Jacobian are both scaled by sqrt(rho').
 * The local Jacobian is J(1x7) * [I6;0] (pose_local_parameterization.cpp:34-40) = the first six columns. */
static inline void residual_block(const oracle_problem* p, const double plane[4], const double pt[3], double scale,
                                  const double pose7[7], double* cost, double* r_out, double* j6_out) {
  double r, j7[7];
  oracle_factor_evaluate(plane, pt, scale, pose7, &r, j6_out ? j7 : NULL);
  if (p->use_loss) {
    double rho[3];
    cauchy_loss(p->cauchy_a * scale, r * r, rho);
    *cost = 0.5 * rho[0];
    const double sq = sqrt(rho[1]);
    if (r_out) *r_out = r * sq;
    if (j6_out) for (int c = 0; c < 6; ++c) j6_out[c] = j7[c] * sq;
  } else {
    *cost = 0.5 * r * r;
    if (r_out) *r_out = r;
    if (j6_out) for (int c = 0; c < 6; ++c) j6_out[c] = j7[c];
  }
}

static inline int frame_has_edges(const oracle_problem* p, int64_t f) {
  return p->edge_points != NULL && p->offsets[f + 1] > p->offsets[f];
}

int64_t oracle_num_residuals(const oracle_problem* p) {
  int64_t n = p->offsets[p->n_frames];
  if (p->edge_points)
    for (int64_t f = 0; f < p->n_frames; ++f) n += frame_has_edges(p, f) ? 2 : 0;
  return n;
}

/* Row index of the first residual of every frame, in the reference's AddResidualBlock order
 * (LaseCamCalCeres.cpp:241-294: a frame's points, then its two edge residuals). */
static int64_t* residual_row_starts(const oracle_problem* p) {
  int64_t* rows = (int64_t*)malloc(sizeof(int64_t) * (size_t)(p->n_frames + 1));
  int64_t r = 0;
  for (int64_t f = 0; f < p->n_frames; ++f) {
    rows[f] = r;
    r += (p->offsets[f + 1] - p->offsets[f]) + (frame_has_edges(p, f) ? 2 : 0);
  }
  rows[p->n_frames] = r;
  return rows;
}

static int resolve_threads(int num_threads) {
#ifdef _OPENMP
  if (num_threads <= 0) return omp_get_max_threads();
  return num_threads;
#else
  (void)num_threads;
  return 1;
#endif
}

/* Problem assembly + one Ceres evaluation (LaseCamCalCeres.cpp:222-295 and Ceres program_evaluator.h). */
int oracle_evaluate(const oracle_problem* p, const double pose7[7], double* cost, double* residuals,
                    double* jacobian, double* gradient, int num_threads) {
  const int nt = resolve_threads(num_threads);
  int64_t* rows = residual_row_starts(p);
  double* part = (double*)calloc((size_t)nt * 8, sizeof(double)); /* per thread: cost + 6 gradient */
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double c_acc = 0.0, g_acc[6] = {0, 0, 0, 0, 0, 0};
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t f = 0; f < p->n_frames; ++f) {
      const int64_t b = p->offsets[f], e = p->offsets[f + 1];
      if (e <= b) continue;
      double plane[4];
      oracle_frame_plane(p->frame_pose + 7 * f, plane);
      double scale = (double)(e - b); /* :239-240 */
      scale = 1. / sqrt(scale);
      int64_t row = rows[f];
      for (int64_t j = b; j < e; ++j, ++row) {
        double c, r, j6[6];
        residual_block(p, plane, p->points + 3 * j, scale, pose7, &c, &r, (jacobian || gradient) ? j6 : NULL);
        c_acc += c;
        if (residuals) residuals[row] = r;
        if (jacobian) memcpy(jacobian + 6 * row, j6, sizeof(j6));
        if (gradient) for (int k = 0; k < 6; ++k) g_acc[k] += j6[k] * r;
      }
      if (frame_has_edges(p, f)) {
        double pi[2][4];
        oracle_edge_planes(p->frame_pose + 7 * f, pi[0], pi[1]);
        for (int k2 = 0; k2 < 2; ++k2, ++row) {
          double c, r, j6[6];
          residual_block(p, pi[k2], p->edge_points + 6 * f + 3 * k2, scale, pose7, &c, &r,
                         (jacobian || gradient) ? j6 : NULL);
          c_acc += c;
          if (residuals) residuals[row] = r;
          if (jacobian) memcpy(jacobian + 6 * row, j6, sizeof(j6));
          if (gradient) for (int k = 0; k < 6; ++k) g_acc[k] += j6[k] * r;
        }
      }
    }
    part[tid * 8] = c_acc;
    for (int k = 0; k < 6; ++k) part[tid * 8 + 1 + k] = g_acc[k];
  }
  double c = 0.0, g[6] = {0, 0, 0, 0, 0, 0};
  for (int t = 0; t < nt; ++t) {
    c += part[t * 8];
    for (int k = 0; k < 6; ++k) g[k] += part[t * 8 + 1 + k];
  }
  if (cost) *cost = c;
  if (gradient) memcpy(gradient, g, sizeof(g));
  free(part);
  free(rows);
  return 0;
}

/* Streaming variant of the same sums: H = J~^T J~, g = J~^T r~, cost = 1/2 sum rho, with R hoisted out of
 * the loop.  Mathematically identical to forming oracle_evaluate()'s J and taking J^T J. */
int oracle_evaluate_normal(const oracle_problem* p, const double pose7[7], double* cost, double* H36, double* g6,
                           int num_threads) {
  const int nt = resolve_threads(num_threads);
  const int want_jac = (H36 != NULL) || (g6 != NULL);
  double R[9];
  oracle_quat_to_rot(pose7 + 3, R);
  const double t0 = pose7[0], t1 = pose7[1], t2 = pose7[2];
  double* part = (double*)calloc((size_t)nt * 28, sizeof(double));
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    double acc[28];
    for (int k = 0; k < 28; ++k) acc[k] = 0.0;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t f = 0; f < p->n_frames; ++f) {
      const int64_t b = p->offsets[f], e = p->offsets[f + 1];
      if (e <= b) continue;
      double planes[3][4];
      oracle_frame_plane(p->frame_pose + 7 * f, planes[0]);
      const int has_edges = frame_has_edges(p, f);
      if (has_edges) oracle_edge_planes(p->frame_pose + 7 * f, planes[1], planes[2]);
      const double scale = 1. / sqrt((double)(e - b));
      const double a = p->cauchy_a * scale, bb = a * a, cc = 1.0 / bb;
      const int64_t n_items = (e - b) + (has_edges ? 2 : 0);
      for (int64_t it = 0; it < n_items; ++it) {
        const double* pl;
        const double* pt;
        if (it < e - b) { pl = planes[0]; pt = p->points + 3 * (b + it); }
        else { pl = planes[1 + (it - (e - b))]; pt = p->edge_points + 6 * f + 3 * (it - (e - b)); }
        /* m = R^T n (so that n^T R p = m.p and n^T(-R [p]x) = (p x m)^T) */
        const double m0 = R[0] * pl[0] + R[3] * pl[1] + R[6] * pl[2];
        const double m1 = R[1] * pl[0] + R[4] * pl[1] + R[7] * pl[2];
        const double m2 = R[2] * pl[0] + R[5] * pl[1] + R[8] * pl[2];
        const double e_raw = (m0 * pt[0] + m1 * pt[1] + m2 * pt[2]) + (pl[0] * t0 + pl[1] * t1 + pl[2] * t2) + pl[3];
        const double r = scale * e_raw;
        double w = 1.0, c;
        if (p->use_loss) {
          const double sum = 1.0 + (r * r) * cc;
          const double inv = 1.0 / sum;
          c = 0.5 * bb * log(sum);
          w = inv > DBL_MIN ? inv : DBL_MIN;
        } else {
          c = 0.5 * r * r;
        }
        acc[27] += c;
        if (want_jac) {
          double J[6];
          J[0] = scale * pl[0]; J[1] = scale * pl[1]; J[2] = scale * pl[2];
          J[3] = scale * (pt[1] * m2 - pt[2] * m1);
          J[4] = scale * (pt[2] * m0 - pt[0] * m2);
          J[5] = scale * (pt[0] * m1 - pt[1] * m0);
          int k = 0;
          for (int i = 0; i < 6; ++i) {
            const double wi = w * J[i];
            for (int j = i; j < 6; ++j) acc[k++] += wi * J[j];
          }
          const double wr = w * r;
          for (int i = 0; i < 6; ++i) acc[21 + i] += wr * J[i];
        }
      }
    }
    memcpy(part + tid * 28, acc, sizeof(acc));
  }
  double tot[28];
  for (int k = 0; k < 28; ++k) tot[k] = 0.0;
  for (int t = 0; t < nt; ++t)
    for (int k = 0; k < 28; ++k) tot[k] += part[t * 28 + k];
  free(part);
  if (H36) {
    int k = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j) { H36[i * 6 + j] = tot[k]; H36[j * 6 + i] = tot[k]; ++k; }
  }
  if (g6) for (int i = 0; i < 6; ++i) g6[i] = tot[21 + i];
  if (cost) *cost = tot[27];
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Dense helpers                                                                                          */
/* ------------------------------------------------------------------------------------------------------ */

/* In-place Householder triangularisation of the rows x 6 row-major A (rows >= 6) and of the rhs b: on exit the
 * top 6 x 6 of A is R and b[0..5] is (Q^T b)[0..5].  A column that is exactly zero from the diagonal down needs no
 * reflection (R_kk = 0, as in Eigen's HouseholderQR, where tau = 0); that is not a failure here: in the blocked use
 * below, a Jacobian column that is identically zero (the reference's "ONLY pitch" boards, main/calibr_simulation.cpp:50:
 * every n_y == 0) becomes non-zero once the LM diagonal rows are stacked under it.  Always returns 0. */
static int hh_factor6(double* A, double* b, int64_t rows) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double nrm2 = 0.0;
    for (int64_t i = k; i < rows; ++i) nrm2 += A[i * n + k] * A[i * n + k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0.0)) { A[(int64_t)k * n + k] = 0.0; continue; }
    const double akk = A[(int64_t)k * n + k];
    const double alpha = akk > 0.0 ? -nrm : nrm;
    /* v = x - alpha e1 (stored in column k below the diagonal, v0 separately); beta = 2 / v^T v */
    const double v0 = akk - alpha;
    const double vtv = nrm2 - akk * akk + v0 * v0;
    if (vtv > 0.0) {
      const double beta = 2.0 / vtv;
      /* one pass for all dot products v^T [A(:,k+1..5) b], one pass for the rank-1 update */
      double s[7];
      for (int j = k + 1; j < n; ++j) s[j] = v0 * A[(int64_t)k * n + j];
      s[6] = v0 * b[k];
      for (int64_t i = k + 1; i < rows; ++i) {
        const double vi = A[i * n + k];
        for (int j = k + 1; j < n; ++j) s[j] += vi * A[i * n + j];
        s[6] += vi * b[i];
      }
      for (int j = k + 1; j < n; ++j) { s[j] *= beta; A[(int64_t)k * n + j] -= s[j] * v0; }
      s[6] *= beta;
      b[k] -= s[6] * v0;
      for (int64_t i = k + 1; i < rows; ++i) {
        const double vi = A[i * n + k];
        for (int j = k + 1; j < n; ++j) A[i * n + j] -= s[j] * vi;
        b[i] -= s[6] * vi;
      }
    }
    A[(int64_t)k * n + k] = alpha;
  }
  return 0;
}

/* Least squares min ||A y - b|| by Householder QR of the rows x 6 row-major A (Ceres: dense_qr_solver.cc ->
 * Eigen householderQr().solve()).  A and b are overwritten.  Tall matrices are processed as a sequence of
 * cache-resident row blocks stacked under the running 6 x 6 triangular factor (a Householder QR of the same matrix,
 * organised so that it streams A once -- this keeps the timed CPU baseline honest).  Returns 0 on success. */
#define HH_BLOCK 2048
static int householder_ls6(double* A, double* b, int64_t rows, double y[6]) {
  const int n = 6;
  double* R = A;
  double* qtb = b;
  double W[(HH_BLOCK + 6) * 6], wb[HH_BLOCK + 6];
  if (rows > HH_BLOCK + 6) {
    int have = 0;
    for (int64_t r0 = 0; r0 < rows; r0 += HH_BLOCK) {
      const int64_t nb = (rows - r0 < HH_BLOCK) ? rows - r0 : HH_BLOCK;
      memcpy(W + have * n, A + r0 * n, sizeof(double) * (size_t)nb * n);
      memcpy(wb + have, b + r0, sizeof(double) * (size_t)nb);
      if (hh_factor6(W, wb, have + nb)) return 1;
      for (int i = 1; i < n; ++i)
        for (int j = 0; j < i; ++j) W[i * n + j] = 0.0; /* keep only R for the next stack */
      have = n;
    }
    R = W;
    qtb = wb;
  } else {
    if (hh_factor6(A, b, rows)) return 1;
  }
  for (int k = n - 1; k >= 0; --k) {
    double sacc = qtb[k];
    for (int j = k + 1; j < n; ++j) sacc -= R[k * n + j] * y[j];
    y[k] = sacc / R[k * n + k];
  }
  return 0;
}

/* Cholesky solve of an SPD n x n system (row-major).  Returns 0 on success. */
static int cholesky_solve(const double* A, const double* b, int n, double* x) {
  double L[81];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      if (i == j) {
        if (!(s > 0.0)) return 1;
        L[i * n + i] = sqrt(s);
      } else {
        L[i * n + j] = s / L[j * n + j];
      }
    }
  double z[9];
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * z[k];
    z[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * x[k];
    x[i] = s / L[i * n + i];
  }
  return 0;
}

/* Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9): A = V diag(w) V^T. */
static void jacobi_eig(const double* Ain, int n, double* w, double* V) {
  double A[81];
  memcpy(A, Ain, sizeof(double) * (size_t)(n * n));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[i * n + i] * A[i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
    }
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; ++i) w[i] = A[i * n + i];
}

/* Singular values (descending) of a symmetric matrix = |eigenvalues| (Eigen::JacobiSVD on H / AtA,
 * LaseCamCalCeres.cpp:162,366). */
void oracle_sym_singular_values(const double* A, int n, double* sv) {
  double w[9], V[81];
  jacobi_eig(A, n, w, V);
  for (int i = 0; i < n; ++i) sv[i] = fabs(w[i]);
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      if (sv[j] > sv[i]) { const double t = sv[i]; sv[i] = sv[j]; sv[j] = t; }
}

/* ------------------------------------------------------------------------------------------------------ */
/* The Ceres trust-region Levenberg-Marquardt solve                                                       */
/* ------------------------------------------------------------------------------------------------------ */

void oracle_default_options(oracle_options* o) {
  /* Ceres: include/ceres/solver.h Solver::Options defaults; the reference overrides only linear_solver_type
   * and max_num_iterations (LaseCamCalCeres.cpp:302-304). */
  o->max_num_iterations = 100;
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->function_tolerance = 1e-6;
  o->gradient_tolerance = 1e-10;
  o->parameter_tolerance = 1e-8;
  o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1;
  o->linear_solver = 0;
  o->num_threads = 1;
}

static double norm7(const double* a) {
  double s = 0.0;
  for (int i = 0; i < 7; ++i) s += a[i] * a[i];
  return sqrt(s);
}

/* Ceres: trust_region_minimizer.cc EvaluateGradientAndJacobian: |x - Plus(x, -g)| in max norm. */
static double gradient_max_norm(const double x[7], const double g[6]) {
  double ng[6], xp[7], m = 0.0;
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  oracle_pose_plus(x, ng, xp);
  for (int i = 0; i < 7; ++i) { const double d = fabs(x[i] - xp[i]); if (d > m) m = d; }
  return m;
}

typedef struct {
  /* Ceres-shaped state (linear_solver == 0) */
  double* residuals;
  double* jacobian; /* scaled by the jacobi scaling, as Ceres keeps it */
  double* qr_A;
  double* qr_b;
  /* normal-equation state (linear_solver == 1), unscaled */
  double H[36], g[6];
} lm_state;

static void record(oracle_iteration* trace, int cap, int* n, const oracle_iteration* it) {
  if (trace && *n < cap) trace[*n] = *it;
  (*n)++;
}

int oracle_solve(const oracle_problem* p, double pose7[7], const oracle_options* opt, oracle_summary* summary,
                 oracle_iteration* trace, int trace_cap) {
  const int64_t R = oracle_num_residuals(p);
  const int qr = (opt->linear_solver == 0);
  lm_state st;
  memset(&st, 0, sizeof(st));
  if (qr) {
    st.residuals = (double*)malloc(sizeof(double) * (size_t)(R > 0 ? R : 1));
    st.jacobian = (double*)malloc(sizeof(double) * (size_t)(R > 0 ? R : 1) * 6);
    st.qr_A = (double*)malloc(sizeof(double) * (size_t)(R + 6) * 6);
    st.qr_b = (double*)malloc(sizeof(double) * (size_t)(R + 6));
  }
  oracle_summary sm;
  memset(&sm, 0, sizeof(sm));
  int n_trace = 0;

  double x[7], cand[7], x_cost, cand_cost, gradient[6], scale[6], diag[6], colnorm2[6];
  memcpy(x, pose7, sizeof(x));
  double x_norm = norm7(x);
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  int reuse_diagonal = 0, num_invalid = 0;

  /* ---- evaluation with Jacobian + jacobi scaling (Ceres: EvaluateGradientAndJacobian) ---- */
#define EVAL_WITH_JACOBIAN(first)                                                                           \
  do {                                                                                                      \
    if (qr) {                                                                                               \
      oracle_evaluate(p, x, &x_cost, st.residuals, st.jacobian, gradient, opt->num_threads);                \
      for (int k = 0; k < 6; ++k) colnorm2[k] = 0.0;                                                        \
      if (first) {                                                                                          \
        for (int64_t i = 0; i < R; ++i)                                                                     \
          for (int k = 0; k < 6; ++k) colnorm2[k] += st.jacobian[i * 6 + k] * st.jacobian[i * 6 + k];       \
        for (int k = 0; k < 6; ++k) scale[k] = opt->jacobi_scaling ? 1.0 / (1.0 + sqrt(colnorm2[k])) : 1.0; \
      }                                                                                                     \
      for (int k = 0; k < 6; ++k) colnorm2[k] = 0.0;                                                        \
      for (int64_t i = 0; i < R; ++i)                                                                       \
        for (int k = 0; k < 6; ++k) {                                                                       \
          st.jacobian[i * 6 + k] *= scale[k];                                                               \
          colnorm2[k] += st.jacobian[i * 6 + k] * st.jacobian[i * 6 + k];                                   \
        }                                                                                                   \
    } else {                                                                                                \
      oracle_evaluate_normal(p, x, &x_cost, st.H, st.g, opt->num_threads);                                  \
      memcpy(gradient, st.g, sizeof(gradient));                                                             \
      if (first)                                                                                            \
        for (int k = 0; k < 6; ++k) scale[k] = opt->jacobi_scaling ? 1.0 / (1.0 + sqrt(st.H[k * 6 + k])) : 1.0; \
      for (int k = 0; k < 6; ++k) colnorm2[k] = scale[k] * scale[k] * st.H[k * 6 + k];                      \
    }                                                                                                       \
    sm.num_residual_evaluations++;                                                                          \
    sm.num_jacobian_evaluations++;                                                                          \
  } while (0)

  /* ---- iteration 0 (Ceres: TrustRegionMinimizer::IterationZero) ---- */
  oracle_iteration it;
  memset(&it, 0, sizeof(it));
  EVAL_WITH_JACOBIAN(1);
  if (!isfinite(x_cost)) {
    /* Ceres: residual_block.cc IsArrayValid -> "Residual and Jacobian evaluation failed" -> FAILURE */
    sm.termination = ORACLE_TERM_FAILURE;
    sm.initial_cost = sm.final_cost = x_cost;
    if (summary) *summary = sm;
    if (qr) { free(st.residuals); free(st.jacobian); free(st.qr_A); free(st.qr_b); }
    return 0;
  }
  it.iteration = 0;
  it.cost = x_cost;
  it.gradient_max_norm = gradient_max_norm(x, gradient);
  it.step_is_valid = 1;
  it.step_is_successful = 1;
  sm.initial_cost = x_cost;
  sm.termination = 0;

  for (;;) {
    /* ---- Ceres: FinalizeIterationAndCheckIfMinimizerCanContinue ---- */
    if (it.step_is_successful) {
      sm.num_successful_steps++;
      memcpy(pose7, x, sizeof(x)); /* monotonic steps: every successful x is the new minimum */
    } else {
      sm.num_unsuccessful_steps++;
    }
    it.trust_region_radius = radius;
    record(trace, trace_cap, &n_trace, &it);
    if (it.iteration >= opt->max_num_iterations) { sm.termination = ORACLE_TERM_NO_CONVERGENCE; break; }
    if (it.step_is_successful && it.gradient_max_norm <= opt->gradient_tolerance) {
      sm.termination = ORACLE_TERM_CONVERGENCE_GRADIENT;
      break;
    }
    if (!(radius > opt->min_trust_region_radius)) { sm.termination = ORACLE_TERM_CONVERGENCE_MIN_RADIUS; break; }

    oracle_iteration prev = it;
    memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1;

    /* ---- Ceres: LevenbergMarquardtStrategy::ComputeStep ---- */
    if (!reuse_diagonal)
      for (int k = 0; k < 6; ++k) {
        double d = colnorm2[k];
        d = d > opt->min_lm_diagonal ? d : opt->min_lm_diagonal;
        d = d < opt->max_lm_diagonal ? d : opt->max_lm_diagonal;
        diag[k] = d;
      }
    double lm_diag[6], step[6];
    for (int k = 0; k < 6; ++k) lm_diag[k] = sqrt(diag[k] / radius);
    int solver_failed;
    if (qr) {
      /* Ceres: DenseQRSolver -- append diag(D) rows to J, rhs = [r; 0], solve J y = r, step = -y */
      memcpy(st.qr_A, st.jacobian, sizeof(double) * (size_t)R * 6);
      memset(st.qr_A + R * 6, 0, sizeof(double) * 36);
      for (int k = 0; k < 6; ++k) st.qr_A[(R + k) * 6 + k] = lm_diag[k];
      memcpy(st.qr_b, st.residuals, sizeof(double) * (size_t)R);
      memset(st.qr_b + R, 0, sizeof(double) * 6);
      solver_failed = householder_ls6(st.qr_A, st.qr_b, R + 6, step);
    } else {
      double Hs[36], gs[6];
      for (int i = 0; i < 6; ++i) {
        gs[i] = scale[i] * st.g[i];
        for (int j = 0; j < 6; ++j) Hs[i * 6 + j] = scale[i] * scale[j] * st.H[i * 6 + j];
        Hs[i * 6 + i] += lm_diag[i] * lm_diag[i];
      }
      solver_failed = cholesky_solve(Hs, gs, 6, step);
    }
    reuse_diagonal = 1;
    for (int k = 0; k < 6; ++k) {
      if (!isfinite(step[k])) solver_failed = 1;
      step[k] = -step[k];
    }

    /* ---- Ceres: TrustRegionMinimizer::ComputeTrustRegionStep: model cost change ---- */
    double model_cost_change = 0.0;
    if (!solver_failed) {
      if (qr) {
        for (int64_t i = 0; i < R; ++i) {
          double mr = 0.0;
          for (int k = 0; k < 6; ++k) mr += st.jacobian[i * 6 + k] * step[k];
          model_cost_change -= mr * (st.residuals[i] + mr / 2.0);
        }
      } else {
        /* -(J s)^T (r + J s / 2) = -g_s.s - 1/2 s^T H_s s */
        double gs_s = 0.0, sHs = 0.0;
        for (int i = 0; i < 6; ++i) {
          gs_s += scale[i] * st.g[i] * step[i];
          for (int j = 0; j < 6; ++j) sHs += step[i] * scale[i] * st.H[i * 6 + j] * scale[j] * step[j];
        }
        model_cost_change = -gs_s - 0.5 * sHs;
      }
      it.step_is_valid = model_cost_change > 0.0;
    }
    if (!it.step_is_valid) {
      /* ---- Ceres: HandleInvalidStep ---- */
      if (++num_invalid >= opt->max_num_consecutive_invalid_steps) { sm.termination = ORACLE_TERM_FAILURE; break; }
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = 1;
      it.cost = x_cost;
      it.gradient_max_norm = prev.gradient_max_norm;
      continue;
    }
    num_invalid = 0;
    double delta[6];
    for (int k = 0; k < 6; ++k) delta[k] = step[k] * scale[k];

    /* ---- Ceres: ComputeCandidatePointAndEvaluateCost ---- */
    oracle_pose_plus(x, delta, cand);
    if (qr) oracle_evaluate(p, cand, &cand_cost, NULL, NULL, NULL, opt->num_threads);
    else oracle_evaluate_normal(p, cand, &cand_cost, NULL, NULL, opt->num_threads);
    sm.num_residual_evaluations++;
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX; /* Ceres: "treating it as a step with infinite cost" */

    /* ---- Ceres: ParameterToleranceReached ---- */
    {
      double d[7];
      for (int i = 0; i < 7; ++i) d[i] = x[i] - cand[i];
      it.step_norm = norm7(d);
    }
    it.cost_change = x_cost - cand_cost;
    it.cost = cand_cost;
    if (it.step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      sm.termination = ORACLE_TERM_CONVERGENCE_PARAMETER;
      it.trust_region_radius = radius;
      record(trace, trace_cap, &n_trace, &it);
      break;
    }
    /* ---- Ceres: FunctionToleranceReached ---- */
    if (fabs(it.cost_change) <= opt->function_tolerance * x_cost) {
      sm.termination = ORACLE_TERM_CONVERGENCE_FUNCTION;
      it.trust_region_radius = radius;
      record(trace, trace_cap, &n_trace, &it);
      break;
    }
    /* ---- Ceres: IsStepSuccessful (monotonic: StepQuality = cost change / model cost change) ---- */
    it.relative_decrease = it.cost_change / model_cost_change;
    if (it.relative_decrease > opt->min_relative_decrease) {
      /* ---- Ceres: HandleSuccessfulStep + LevenbergMarquardtStrategy::StepAccepted ---- */
      memcpy(x, cand, sizeof(x));
      x_norm = norm7(x);
      EVAL_WITH_JACOBIAN(0);
      it.cost = x_cost;
      it.gradient_max_norm = gradient_max_norm(x, gradient);
      it.step_is_successful = 1;
      {
        const double q = 2.0 * it.relative_decrease - 1.0;
        double den = 1.0 - q * q * q;
        if (den < 1.0 / 3.0) den = 1.0 / 3.0;
        radius = radius / den;
        if (radius > opt->max_trust_region_radius) radius = opt->max_trust_region_radius;
      }
      decrease_factor = 2.0;
      reuse_diagonal = 0;
    } else {
      /* ---- Ceres: HandleUnsuccessfulStep + StepRejected ---- */
      it.step_is_successful = 0;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = 1;
    }
  }
#undef EVAL_WITH_JACOBIAN

  sm.num_iterations = n_trace;
  sm.final_cost = x_cost;
  if (summary) *summary = sm;
  if (qr) { free(st.residuals); free(st.jacobian); free(st.qr_A); free(st.qr_b); }
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Scan preparation: src/utilities.cpp:181-215 and src/selectScanPoints.cpp:17-190                        */
/* ------------------------------------------------------------------------------------------------------ */
void oracle_scan_to_points(const float* ranges, int64_t n, double angle_min, double angle_increment, double range_min,
                           double* points) {
  for (int64_t i = 0; i < n; ++i) {
    const double ang = angle_min + (double)i * angle_increment;
    const float range = ranges[i];
    if (range < 30.0 && range >= range_min) { /* range_cutoff = 30 (:205) */
      points[3 * i] = (double)range * cos(ang);
      points[3 * i + 1] = (double)range * sin(ang);
    } else {
      points[3 * i] = 1000.0;
      points[3 * i + 1] = 1000.0;
    }
    points[3 * i + 2] = 0.0;
  }
}

static inline double norm_xy(const double* points, int64_t i) {
  return sqrt(points[3 * i] * points[3 * i] + points[3 * i + 1] * points[3 * i + 1]);
}

int oracle_auto_get_line_pts(const double* points, int64_t n, int64_t* start, int64_t* end) {
  if (n <= 0) return 0;
  const int64_t id = n / 2;
  const int64_t delta = (int64_t)(80 / 0.3); /* :41 */
  const int64_t id_left = id + delta < n - 1 ? id + delta : n - 1;
  const int64_t id_right = id - delta > 0 ? id - delta : 0;
  const double dist_thre = 0.05, range_max = 100;
  const int skip = 3;
  int64_t best_start = -1, best_end = -1, best_cnt = -1;
  int64_t cur = id_right, next = cur + skip, seg_start = 0, seg_end = 0;
  int new_seg = 1;
  for (int64_t i = id_right; i < id_left - skip; i += skip) { /* :58 */
    if (new_seg) { seg_start = cur; seg_end = next; new_seg = 0; }
    const double d1 = norm_xy(points, cur), d2 = norm_xy(points, next);
    if (d1 < range_max && d2 < range_max) {
      if (fabs(d1 - d2) < dist_thre) {
        seg_end = next;
      } else {
        new_seg = 1;
        const double dx = points[3 * seg_start] - points[3 * seg_end], dy = points[3 * seg_start + 1] - points[3 * seg_end + 1];
        if (sqrt(dx * dx + dy * dy) > 0.2 && norm_xy(points, seg_start) < 2 && norm_xy(points, seg_end) < 2 &&
            seg_end - seg_start > 50) { /* :79-82 */
          /* boundary extension (:103-129); the reference indexes with .at() and would throw outside [0,n): skipped here */
          int64_t s = seg_start, e = seg_end;
          for (int j = 1; j < 4; ++j) {
            const int64_t bp = seg_end + j;
            if (bp < n && fabs(norm_xy(points, seg_end) - norm_xy(points, bp)) < dist_thre) e = bp;
          }
          for (int j = -1; j > -4; --j) {
            const int64_t bp = seg_start + j;
            if (bp >= 0 && fabs(norm_xy(points, seg_start) - norm_xy(points, bp)) < dist_thre) s = bp;
          }
          if (e - s > best_cnt) { best_cnt = e - s; best_start = s; best_end = e; } /* :136-146, first maximum wins */
        }
      }
      cur = next;
      next += skip;
    } else {
      if (d1 > range_max) cur = next;
      next += skip;
    }
  }
  if (best_cnt < 0) return 0;
  *start = best_start;
  *end = best_end;
  return 1;
}

/* ------------------------------------------------------------------------------------------------------ */
/* LineFittingCeres, LaseCamCalCeres.cpp:385-433                                                          */
/* ------------------------------------------------------------------------------------------------------ */

/* min ||A y - b|| for a rows x 2 row-major A by Householder QR (in place). */
static int householder_ls2(double* A, double* b, int64_t rows, double y[2]) {
  for (int k = 0; k < 2; ++k) {
    double nrm2 = 0.0;
    for (int64_t i = k; i < rows; ++i) nrm2 += A[i * 2 + k] * A[i * 2 + k];
    const double nrm = sqrt(nrm2);
    if (!(nrm > 0.0)) return 1;
    const double akk = A[k * 2 + k];
    const double alpha = akk > 0.0 ? -nrm : nrm;
    const double v0 = akk - alpha;
    const double vtv = nrm2 - akk * akk + v0 * v0;
    if (vtv > 0.0) {
      const double beta = 2.0 / vtv;
      for (int j = k + 1; j <= 2; ++j) { /* j == 2 is the right-hand side */
        double s = v0 * (j < 2 ? A[k * 2 + j] : b[k]);
        for (int64_t i = k + 1; i < rows; ++i) s += A[i * 2 + k] * (j < 2 ? A[i * 2 + j] : b[i]);
        s *= beta;
        if (j < 2) {
          A[k * 2 + j] -= s * v0;
          for (int64_t i = k + 1; i < rows; ++i) A[i * 2 + j] -= s * A[i * 2 + k];
        } else {
          b[k] -= s * v0;
          for (int64_t i = k + 1; i < rows; ++i) b[i] -= s * A[i * 2 + k];
        }
      }
    }
    A[k * 2 + k] = alpha;
  }
  y[1] = b[1] / A[3];
  y[0] = (b[0] - A[1] * y[1]) / A[0];
  return 0;
}

/* One Ceres evaluation of the line problem: residual m0 x + m1 y + 1 (:391), AutoDiff Jacobian (x, y), CauchyLoss(0.05)
 * (:416) through the Corrector's simple branch.  r and J may be NULL (cost only). */
static double line_evaluate(const double* pts, int64_t n, const double m[2], double* r, double* J, double g[2]) {
  const double a = 0.05, b = a * a, c = 1.0 / b;
  double cost = 0.0;
  if (g) g[0] = g[1] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double x = pts[3 * i], y = pts[3 * i + 1];
    const double res = m[0] * x + m[1] * y + 1.;
    const double sum = 1.0 + (res * res) * c;
    const double inv = 1.0 / sum;
    cost += 0.5 * b * log(sum);
    if (r) {
      const double sq = sqrt(inv > DBL_MIN ? inv : DBL_MIN);
      r[i] = res * sq;
      J[2 * i] = x * sq;
      J[2 * i + 1] = y * sq;
      if (g) { g[0] += J[2 * i] * r[i]; g[1] += J[2 * i + 1] * r[i]; }
    }
  }
  return cost;
}

/* The same Ceres trust-region loop as oracle_solve (see there for the line-by-line citations), for 2 Euclidean
 * parameters: Plus is x + delta, the gradient norm is |g|_inf, the parameter norm is the 2-norm of (m0, m1). */
int oracle_line_fit(const double* points, int64_t n, double line[2], int max_num_iterations, oracle_summary* summary,
                    oracle_iteration* trace, int trace_cap) {
  oracle_options opt_s;
  oracle_default_options(&opt_s);
  const oracle_options* opt = &opt_s;
  opt_s.max_num_iterations = max_num_iterations;
  oracle_summary sm;
  memset(&sm, 0, sizeof(sm));
  int n_trace = 0;
  double* res = (double*)malloc(sizeof(double) * (size_t)(n + 2));
  double* J = (double*)malloc(sizeof(double) * (size_t)(n + 2) * 2);
  double* qA = (double*)malloc(sizeof(double) * (size_t)(n + 2) * 2);
  double* qb = (double*)malloc(sizeof(double) * (size_t)(n + 2));
  double x[2] = {line[0], line[1]}, cand[2], g[2], scale[2], diag[2], col2[2];
  double x_cost = line_evaluate(points, n, x, res, J, g);
  sm.num_residual_evaluations = sm.num_jacobian_evaluations = 1;
  if (!isfinite(x_cost)) {
    sm.termination = ORACLE_TERM_FAILURE;
    sm.initial_cost = sm.final_cost = x_cost;
    if (summary) *summary = sm;
    free(res); free(J); free(qA); free(qb);
    return 0;
  }
#define LINE_SCALE_J(first)                                                                     \
  do {                                                                                          \
    if (first) {                                                                                \
      col2[0] = col2[1] = 0.0;                                                                  \
      for (int64_t i = 0; i < n; ++i) { col2[0] += J[2 * i] * J[2 * i]; col2[1] += J[2 * i + 1] * J[2 * i + 1]; } \
      scale[0] = 1.0 / (1.0 + sqrt(col2[0]));                                                   \
      scale[1] = 1.0 / (1.0 + sqrt(col2[1]));                                                   \
    }                                                                                           \
    col2[0] = col2[1] = 0.0;                                                                    \
    for (int64_t i = 0; i < n; ++i) {                                                           \
      J[2 * i] *= scale[0]; J[2 * i + 1] *= scale[1];                                           \
      col2[0] += J[2 * i] * J[2 * i]; col2[1] += J[2 * i + 1] * J[2 * i + 1];                   \
    }                                                                                           \
  } while (0)
  LINE_SCALE_J(1);
  double x_norm = sqrt(x[0] * x[0] + x[1] * x[1]);
  double radius = opt->initial_trust_region_radius, decrease_factor = 2.0;
  int reuse_diagonal = 0, num_invalid = 0;
  oracle_iteration it;
  memset(&it, 0, sizeof(it));
  it.cost = x_cost;
  it.gradient_max_norm = fmax(fabs(g[0]), fabs(g[1]));
  it.step_is_valid = it.step_is_successful = 1;
  sm.initial_cost = x_cost;
  for (;;) {
    if (it.step_is_successful) { sm.num_successful_steps++; line[0] = x[0]; line[1] = x[1]; }
    else sm.num_unsuccessful_steps++;
    it.trust_region_radius = radius;
    record(trace, trace_cap, &n_trace, &it);
    if (it.iteration >= opt->max_num_iterations) { sm.termination = ORACLE_TERM_NO_CONVERGENCE; break; }
    if (it.step_is_successful && it.gradient_max_norm <= opt->gradient_tolerance) { sm.termination = ORACLE_TERM_CONVERGENCE_GRADIENT; break; }
    if (!(radius > opt->min_trust_region_radius)) { sm.termination = ORACLE_TERM_CONVERGENCE_MIN_RADIUS; break; }
    oracle_iteration prev = it;
    memset(&it, 0, sizeof(it));
    it.iteration = prev.iteration + 1;
    if (!reuse_diagonal)
      for (int k = 0; k < 2; ++k) {
        double d = col2[k];
        d = d > opt->min_lm_diagonal ? d : opt->min_lm_diagonal;
        d = d < opt->max_lm_diagonal ? d : opt->max_lm_diagonal;
        diag[k] = d;
      }
    double step[2];
    memcpy(qA, J, sizeof(double) * (size_t)n * 2);
    qA[2 * n] = sqrt(diag[0] / radius); qA[2 * n + 1] = 0.0;
    qA[2 * n + 2] = 0.0; qA[2 * n + 3] = sqrt(diag[1] / radius);
    memcpy(qb, res, sizeof(double) * (size_t)n);
    qb[n] = qb[n + 1] = 0.0;
    int failed = householder_ls2(qA, qb, n + 2, step);
    reuse_diagonal = 1;
    for (int k = 0; k < 2; ++k) { if (!isfinite(step[k])) failed = 1; step[k] = -step[k]; }
    double mcc = 0.0;
    if (!failed) {
      for (int64_t i = 0; i < n; ++i) {
        const double mr = J[2 * i] * step[0] + J[2 * i + 1] * step[1];
        mcc -= mr * (res[i] + mr / 2.0);
      }
      it.step_is_valid = mcc > 0.0;
    }
    if (!it.step_is_valid) {
      if (++num_invalid >= opt->max_num_consecutive_invalid_steps) { sm.termination = ORACLE_TERM_FAILURE; break; }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
      it.cost = x_cost; it.gradient_max_norm = prev.gradient_max_norm;
      continue;
    }
    num_invalid = 0;
    cand[0] = x[0] + step[0] * scale[0];
    cand[1] = x[1] + step[1] * scale[1];
    double cand_cost = line_evaluate(points, n, cand, NULL, NULL, NULL);
    sm.num_residual_evaluations++;
    if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
    it.step_norm = sqrt((x[0] - cand[0]) * (x[0] - cand[0]) + (x[1] - cand[1]) * (x[1] - cand[1]));
    it.cost_change = x_cost - cand_cost;
    it.cost = cand_cost;
    if (it.step_norm <= opt->parameter_tolerance * (x_norm + opt->parameter_tolerance)) {
      sm.termination = ORACLE_TERM_CONVERGENCE_PARAMETER; it.trust_region_radius = radius; record(trace, trace_cap, &n_trace, &it); break;
    }
    if (fabs(it.cost_change) <= opt->function_tolerance * x_cost) {
      sm.termination = ORACLE_TERM_CONVERGENCE_FUNCTION; it.trust_region_radius = radius; record(trace, trace_cap, &n_trace, &it); break;
    }
    it.relative_decrease = it.cost_change / mcc;
    if (it.relative_decrease > opt->min_relative_decrease) {
      x[0] = cand[0]; x[1] = cand[1];
      x_norm = sqrt(x[0] * x[0] + x[1] * x[1]);
      x_cost = line_evaluate(points, n, x, res, J, g);
      sm.num_residual_evaluations++; sm.num_jacobian_evaluations++;
      LINE_SCALE_J(0);
      it.cost = x_cost;
      it.gradient_max_norm = fmax(fabs(g[0]), fabs(g[1]));
      it.step_is_successful = 1;
      const double q = 2.0 * it.relative_decrease - 1.0;
      double den = 1.0 - q * q * q;
      if (den < 1.0 / 3.0) den = 1.0 / 3.0;
      radius /= den;
      if (radius > opt->max_trust_region_radius) radius = opt->max_trust_region_radius;
      decrease_factor = 2.0; reuse_diagonal = 0;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = 1;
    }
  }
#undef LINE_SCALE_J
  sm.num_iterations = n_trace;
  sm.final_cost = x_cost;
  if (summary) *summary = sm;
  free(res); free(J); free(qA); free(qb);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Analysis tail, LaseCamCalCeres.cpp:318-381                                                             */
/* ------------------------------------------------------------------------------------------------------ */
int oracle_information(const oracle_problem* p, const double pose7[7], double* H36, double* b6, double* chi,
                       double* sv6) {
  double H[36], b[6], c = 0.0;
  memset(H, 0, sizeof(H));
  memset(b, 0, sizeof(b));
  for (int64_t f = 0; f < p->n_frames; ++f) {
    const int64_t bg = p->offsets[f], en = p->offsets[f + 1];
    if (en <= bg) continue;
    double plane[4];
    oracle_frame_plane(p->frame_pose + 7 * f, plane);
    const double scale = 1. / sqrt((double)(en - bg));
    for (int64_t j = bg; j < en; ++j) {
      double r, j7[7];
      oracle_factor_evaluate(plane, p->points + 3 * j, scale, pose7, &r, j7);
      for (int i = 0; i < 6; ++i) {
        for (int k = 0; k < 6; ++k) H[i * 6 + k] += j7[i] * j7[k];
        b[i] -= j7[i] * r;
      }
      c += r * r;
    }
  }
  if (H36) memcpy(H36, H, sizeof(H));
  if (b6) memcpy(b6, b, sizeof(b));
  if (chi) *chi = c;
  if (sv6) oracle_sym_singular_values(H, 6, sv6);
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Closed form, LaseCamCalCeres.cpp:112-203                                                               */
/* ------------------------------------------------------------------------------------------------------ */

/* LDLT with symmetric (diagonal) pivoting, the scheme of Eigen's LDLT, for a positive semi-definite 9x9. */
static int ldlt_solve9(const double* Ain, const double* bin, double* x) {
  const int n = 9;
  double A[81], b[9];
  int perm[9];
  memcpy(A, Ain, sizeof(A));
  memcpy(b, bin, sizeof(b));
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(A[i * n + i]) > fabs(A[piv * n + piv])) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; ++j) { const double t = A[k * n + j]; A[k * n + j] = A[piv * n + j]; A[piv * n + j] = t; }
      for (int j = 0; j < n; ++j) { const double t = A[j * n + k]; A[j * n + k] = A[j * n + piv]; A[j * n + piv] = t; }
      { const double t = b[k]; b[k] = b[piv]; b[piv] = t; }
      { const int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t; }
    }
    const double d = A[k * n + k];
    if (d == 0.0) continue;
    for (int i = k + 1; i < n; ++i) {
      const double l = A[i * n + k] / d;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
      A[i * n + k] = l;
    }
  }
  /* forward: L z = b */
  double z[9];
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[i * n + k] * z[k];
    z[i] = s;
  }
  for (int i = 0; i < n; ++i) z[i] = (A[i * n + i] != 0.0) ? z[i] / A[i * n + i] : 0.0;
  double y[9];
  for (int i = n - 1; i >= 0; --i) {
    double s = z[i];
    for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * y[k];
    y[i] = s;
  }
  for (int i = 0; i < n; ++i) x[perm[i]] = y[i];
  return 0;
}

int oracle_closed_form(const oracle_problem* p, double Tlc[16], int* unobservable, double* AtA81, double* Atb9) {
  double AtA[81], Atb[9];
  memset(AtA, 0, sizeof(AtA));
  memset(Atb, 0, sizeof(Atb));
  for (int64_t f = 0; f < p->n_frames; ++f) {
    double plane[4];
    oracle_frame_plane(p->frame_pose + 7 * f, plane);
    for (int64_t j = p->offsets[f]; j < p->offsets[f + 1]; ++j) {
      const double bar[3] = {p->points[3 * j], p->points[3 * j + 1], 1.0}; /* :147 */
      double Ai[9];
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) Ai[c * 3 + r] = plane[r] * bar[c]; /* :150-152 */
      const double bi = -plane[3];                                    /* :155 */
      for (int a = 0; a < 9; ++a) {
        for (int b2 = 0; b2 < 9; ++b2) AtA[a * 9 + b2] += Ai[a] * Ai[b2];
        Atb[a] += Ai[a] * bi;
      }
    }
  }
  if (AtA81) memcpy(AtA81, AtA, sizeof(AtA));
  if (Atb9) memcpy(Atb9, Atb, sizeof(Atb));
  double sv[9];
  oracle_sym_singular_values(AtA, 9, sv);
  int unobs = 0;
  for (int i = 0; i < 9; ++i)
    if (sv[i] < 1e-10) unobs = 1; /* :165-171 */
  if (unobservable) *unobservable = unobs;
  double H[9];
  ldlt_solve9(AtA, Atb, H); /* :181 */
  const double* h1 = H;
  const double* h2 = H + 3;
  const double* h3 = H + 6;
  /* Rcl = [h1 h2 h1xh2] (columns); Rlc = Rcl^T; tlc = -Rlc h3   (:187-192) */
  const double h12[3] = {h1[1] * h2[2] - h1[2] * h2[1], h1[2] * h2[0] - h1[0] * h2[2], h1[0] * h2[1] - h1[1] * h2[0]};
  double Rlc[9] = {h1[0], h1[1], h1[2], h2[0], h2[1], h2[2], h12[0], h12[1], h12[2]};
  double tlc[3];
  for (int r = 0; r < 3; ++r) tlc[r] = -(Rlc[r * 3] * h3[0] + Rlc[r * 3 + 1] * h3[1] + Rlc[r * 3 + 2] * h3[2]);
  /* nearest orthogonal matrix U V^T of Rlc (:195-196) = Rlc (Rlc^T Rlc)^(-1/2), no determinant check */
  double G[9], w[3], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) G[i * 3 + j] = Rlc[0 * 3 + i] * Rlc[0 * 3 + j] + Rlc[1 * 3 + i] * Rlc[1 * 3 + j] + Rlc[2 * 3 + i] * Rlc[2 * 3 + j];
  jacobi_eig(G, 3, w, V);
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += V[i * 3 + k] * (1.0 / sqrt(w[k])) * V[j * 3 + k];
      S[i * 3 + j] = s;
    }
  double Ro[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ro[i * 3 + j] = Rlc[i * 3] * S[0 * 3 + j] + Rlc[i * 3 + 1] * S[1 * 3 + j] + Rlc[i * 3 + 2] * S[2 * 3 + j];
  for (int i = 0; i < 16; ++i) Tlc[i] = 0.0;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Tlc[r * 4 + c] = Ro[r * 3 + c];
    Tlc[r * 4 + 3] = tlc[r];
  }
  Tlc[15] = 1.0;
  return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* Synthetic generator, main/calibr_simulation.cpp:10-108 with a counter-based RNG                        */
/* ------------------------------------------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al., SC'11).  key = seed, counter = (ctr_lo, ctr_hi). */
void oracle_philox4x32(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline double u53(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

/* Stream ids in the top byte of ctr_hi. */
#define GEN_STREAM_POSE ((uint64_t)1 << 56)
#define GEN_STREAM_NOISE ((uint64_t)2 << 56)

static const double GEN_RLC[9] = {0, 0, 1, -1, 0, 0, 0, -1, 0}; /* calibr_simulation.cpp:15-18 */
static const double GEN_TLC[3] = {0.1, 0.2, 0.3};               /* :20 */

void oracle_gen_ground_truth(double Tlc[16], double Tcl_pose7[7]) {
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Tlc[r * 4 + c] = GEN_RLC[r * 3 + c];
    Tlc[r * 4 + 3] = GEN_TLC[r];
  }
  Tlc[12] = Tlc[13] = Tlc[14] = 0.0; Tlc[15] = 1.0;
  /* Tcl = Tlc^-1 = [Rlc^T, -Rlc^T tlc] */
  double Tcl[16];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) Tcl[r * 4 + c] = GEN_RLC[c * 3 + r];
    Tcl[r * 4 + 3] = -(GEN_RLC[0 * 3 + r] * GEN_TLC[0] + GEN_RLC[1 * 3 + r] * GEN_TLC[1] + GEN_RLC[2 * 3 + r] * GEN_TLC[2]);
  }
  Tcl[12] = Tcl[13] = Tcl[14] = 0.0; Tcl[15] = 1.0;
  oracle_T_to_pose7(Tcl, Tcl_pose7);
}

/* Board pose draw (:30-32,42-44,58): yaw,pitch,roll ~ U(-pi/6,pi/6), Rca = Rz(yaw) Ry(pitch) Rx(roll);
 * tca = (U(-3,3), U(-3,3), U(1,5)). */
static void gen_draw_pose(uint64_t seed, int64_t frame, int attempt, double fp[7]) {
  double u[6];
  for (int b = 0; b < 3; ++b) {
    uint32_t o[4];
    oracle_philox4x32(seed, (uint64_t)frame, GEN_STREAM_POSE | ((uint64_t)attempt << 8) | (uint64_t)b, o);
    u[2 * b] = u53(o[0], o[1]);
    u[2 * b + 1] = u53(o[2], o[3]);
  }
  const double lim = M_PI / 6.;
  const double yaw = -lim + 2.0 * lim * u[0], pitch = -lim + 2.0 * lim * u[1], roll = -lim + 2.0 * lim * u[2];
  const double cz = cos(yaw), sz = sin(yaw), cy = cos(pitch), sy = sin(pitch), cx = cos(roll), sx = sin(roll);
  const double R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                       sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                       -sy,     cy * sx,                cy * cx};
  oracle_rot_to_quat(R, fp); /* :60 Eigen::Quaterniond qca(Rca) */
  fp[4] = -3.0 + 6.0 * u[3];
  fp[5] = -3.0 + 6.0 * u[4];
  fp[6] = 1.0 + 4.0 * u[5];
}

/* Board plane in the laser frame (:62-73): Rla = Rlc R(qca), tla = Rlc tca + tlc, n = Rla e_z, d = -n.tla. */
static void gen_plane_laser(const double fp[7], double nl[3], double* dl) {
  double Rca[9], tla[3];
  oracle_quat_to_rot(fp, Rca);
  for (int r = 0; r < 3; ++r) {
    nl[r] = GEN_RLC[r * 3] * Rca[2] + GEN_RLC[r * 3 + 1] * Rca[5] + GEN_RLC[r * 3 + 2] * Rca[8];
    tla[r] = (GEN_RLC[r * 3] * fp[4] + GEN_RLC[r * 3 + 1] * fp[5] + GEN_RLC[r * 3 + 2] * fp[6]) + GEN_TLC[r];
  }
  *dl = -(nl[0] * tla[0] + nl[1] * tla[1] + nl[2] * tla[2]);
}

/* exact-M mode: the valid beams of :83-94 (depth > 0, |x| < 5, |y| < 5, theta in [-pi/2, pi/2)) are exactly the
 * part of the 2-D line nx x + ny y + d = 0 inside the box 0 <= x < 5, |y| < 5, a single segment.  Returns 0 when
 * the (slightly shrunk) segment is shorter than 0.2 m, else the beam-angle window [th_a, th_b]. */
static int gen_window(const double nl[3], double dl, double* th_a, double* th_b) {
  const double rho2 = nl[0] * nl[0] + nl[1] * nl[1];
  if (!(rho2 > 1e-12) || !(fabs(dl) > 1e-9)) return 0;
  const double rho = sqrt(rho2);
  const double ux = -nl[1] / rho, uy = nl[0] / rho;
  const double px = -dl * nl[0] / rho2, py = -dl * nl[1] / rho2;
  const double lim = 5.0 * (1.0 - 1e-3);
  double s0 = -1e30, s1 = 1e30;
  /* clip p + s u to 0 <= x <= lim, -lim <= y <= lim */
  const double lo[2] = {0.0, -lim}, hi[2] = {lim, lim}, pp[2] = {px, py}, uu[2] = {ux, uy};
  for (int a = 0; a < 2; ++a) {
    if (fabs(uu[a]) < 1e-14) {
      if (pp[a] < lo[a] || pp[a] > hi[a]) return 0;
    } else {
      double ta = (lo[a] - pp[a]) / uu[a], tb = (hi[a] - pp[a]) / uu[a];
      if (ta > tb) { const double t = ta; ta = tb; tb = t; }
      if (ta > s0) s0 = ta;
      if (tb < s1) s1 = tb;
    }
  }
  if (!(s1 - s0 >= 0.2)) return 0;
  *th_a = atan2(py + s0 * uy, px + s0 * ux);
  *th_b = atan2(py + s1 * uy, px + s1 * ux);
  return 1;
}

/* Edge points for the boundary residuals: intersection of the board-edge lines p1p2 / p1p3 (board corners of
 * LaseCamCalCeres.cpp:262-268, mapped camera -> laser frame with the ground truth) with the scan plane z_l = 0,
 * so that both edge residuals vanish at ground truth.  Returns 0 if an intersection is degenerate / too far. */
static int gen_edge_points(const double fp[7], double ep[6]) {
  const double orig = 0.0265 + 0.0165;
  const double pm[3][3] = {{-orig, -orig, 0.0}, {0.5 - orig, -orig, 0.0}, {-orig, 0.5 - orig, 0.0}};
  double Rca[9], pl[3][3];
  oracle_quat_to_rot(fp, Rca);
  for (int k = 0; k < 3; ++k) {
    double pc[3];
    for (int r = 0; r < 3; ++r)
      pc[r] = (Rca[r * 3] * pm[k][0] + Rca[r * 3 + 1] * pm[k][1] + Rca[r * 3 + 2] * pm[k][2]) + fp[4 + r];
    for (int r = 0; r < 3; ++r)
      pl[k][r] = (GEN_RLC[r * 3] * pc[0] + GEN_RLC[r * 3 + 1] * pc[1] + GEN_RLC[r * 3 + 2] * pc[2]) + GEN_TLC[r];
  }
  for (int e = 0; e < 2; ++e) {
    const double* a = pl[0];
    const double* b = pl[1 + e];
    const double dz = b[2] - a[2];
    if (!(fabs(dz) > 1e-9)) return 0;
    const double lam = -a[2] / dz;
    if (!(fabs(lam) <= 8.0)) return 0;
    ep[3 * e] = a[0] + lam * (b[0] - a[0]);
    ep[3 * e + 1] = a[1] + lam * (b[1] - a[1]);
    ep[3 * e + 2] = 0.0;
  }
  return 1;
}

/* Accept / redraw rule shared by both passes: exact-M mode needs a usable window, edge mode needs usable
 * edge intersections; up to 64 attempts per frame. */
static void gen_frame_pose(const oracle_gen_desc* g, int64_t f, double fp[7]) {
  for (int attempt = 0; attempt < 64; ++attempt) {
    gen_draw_pose(g->seed, f, attempt, fp);
    int ok = 1;
    if (g->exact_m) {
      double nl[3], dl, a, b;
      gen_plane_laser(fp, nl, &dl);
      ok = gen_window(nl, dl, &a, &b);
    }
    if (ok && g->with_edges) {
      double ep[6];
      ok = gen_edge_points(fp, ep);
    }
    if (ok) return;
  }
}

static inline double gen_noise(const oracle_gen_desc* g, int64_t f, int64_t j) {
  if (!(g->sigma > 0.0)) return 0.0;
  uint32_t o[4];
  oracle_philox4x32(g->seed, (uint64_t)f, GEN_STREAM_NOISE | (uint64_t)j, o);
  const double u1 = u53(o[0], o[1]), u2 = u53(o[2], o[3]);
  return g->sigma * sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * M_PI * u2); /* Box-Muller */
}

/* One beam (:81-94).  Returns 1 if the point is valid. */
static inline int gen_beam(const double nl[3], double dl, double theta, double noise, double p[3]) {
  const double cx = cos(theta), sy = sin(theta);
  double depth = -dl / (cx * nl[0] + sy * nl[1]);
  if (isnan(depth) || depth < 0) return 0;
  depth += noise;
  p[0] = depth * cx; p[1] = depth * sy; p[2] = 0.0;
  return fabs(p[0]) < 5 && fabs(p[1]) < 5;
}

int64_t oracle_gen_frames(const oracle_gen_desc* g, double* frame_pose, int64_t* offsets) {
  int64_t total = 0;
  offsets[0] = 0;
  for (int64_t f = 0; f < g->n_frames; ++f) {
    double* fp = frame_pose + 7 * f;
    gen_frame_pose(g, f, fp);
    int64_t cnt = 0;
    if (g->exact_m) {
      cnt = g->beams;
    } else {
      double nl[3], dl, p[3];
      gen_plane_laser(fp, nl, &dl);
      for (int64_t j = 0; j < g->beams; ++j) {
        const double theta = -M_PI_2 + (double)j * M_PI / (double)g->beams; /* :81 with 180 -> beams */
        cnt += gen_beam(nl, dl, theta, gen_noise(g, f, j), p);
      }
    }
    total += cnt;
    offsets[f + 1] = total;
  }
  return total;
}

int oracle_gen_points(const oracle_gen_desc* g, const double* frame_pose, const int64_t* offsets, double* points,
                      double* edge_points) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t f = 0; f < g->n_frames; ++f) {
    const double* fp = frame_pose + 7 * f;
    double nl[3], dl;
    gen_plane_laser(fp, nl, &dl);
    double* out = points + 3 * offsets[f];
    if (g->exact_m) {
      double a = 0.0, b = 0.0;
      gen_window(nl, dl, &a, &b);
      for (int64_t j = 0; j < g->beams; ++j) {
        const double theta = a + (b - a) * (((double)j + 0.5) / (double)g->beams);
        const double cx = cos(theta), sy = sin(theta);
        const double depth = -dl / (cx * nl[0] + sy * nl[1]) + gen_noise(g, f, j);
        out[3 * j] = depth * cx; out[3 * j + 1] = depth * sy; out[3 * j + 2] = 0.0;
      }
    } else {
      int64_t k = 0;
      for (int64_t j = 0; j < g->beams; ++j) {
        const double theta = -M_PI_2 + (double)j * M_PI / (double)g->beams;
        double p[3];
        if (gen_beam(nl, dl, theta, gen_noise(g, f, j), p)) { memcpy(out + 3 * k, p, sizeof(p)); ++k; }
      }
    }
    if (g->with_edges && edge_points) {
      if (!gen_edge_points(fp, edge_points + 6 * f))
        for (int k = 0; k < 6; ++k) edge_points[6 * f + k] = 0.0;
    }
  }
  return 0;
}

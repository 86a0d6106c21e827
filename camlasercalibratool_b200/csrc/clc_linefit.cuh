// clc_linefit.cuh -- batched LineFittingCeres (reference src/LaseCamCalCeres.cpp:385-433): the per-scan robust fit of
// m0 x + m1 y + 1 = 0 that produces Oberserve::points_on_line in main/calibr_offline.cpp:123-142.
//
// One warp per scan, the whole Ceres loop (<= 10 iterations, CauchyLoss(0.05), DENSE_QR-equivalent 2x2 step) inside the
// kernel: every LM iteration is one pass of the warp over the scan's points (a few hundred, L1/L2 resident after the
// first pass), a shuffle reduction of 5 sums + the cost product, and the same trust-region state machine as the pose
// solve, specialised to 2 Euclidean parameters (Plus is x + delta) and run redundantly by all lanes.
#pragma once

#include "clc_math.cuh"
#include "../../include/clc_b200.h"

namespace clc {

// Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy for two Euclidean parameters.  Same rules, same order of
// tests and same defaults as lm_update (clc_lm.cuh); the candidate is evaluated together with its Jacobian.
struct Lm2 {
  double x[2], cand[2];
  double x_cost, x_norm;
  double H[3], g[2];  // H = (xx, xy, yy) at x, loss-corrected, unscaled
  double scale[2], diag[2];
  double radius, decrease_factor, model_cost_change, initial_cost;
  int done, phase, iteration, num_invalid, reuse_diagonal, sweeps;
};

CLC_HD void lm2_init(Lm2& s, double m0, double m1) {
  s.x[0] = s.cand[0] = m0;
  s.x[1] = s.cand[1] = m1;
  s.x_cost = 0.0;
  s.x_norm = sqrt(m0 * m0 + m1 * m1);
  s.radius = 1e4;
  s.decrease_factor = 2.0;
  s.model_cost_change = 0.0;
  s.initial_cost = 0.0;
  s.done = 0; s.phase = 0; s.iteration = 0; s.num_invalid = 0; s.reuse_diagonal = 0; s.sweeps = 0;
}

// sums = (Hxx, Hxy, Hyy, gx, gy, cost) of the sweep that evaluated s.cand
CLC_HD void lm2_update(Lm2& s, const double* sums, int max_num_iterations) {
  if (s.done) return;
  s.sweeps++;
  const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, min_rel = 1e-3, min_diag = 1e-6, max_diag = 1e32;
  bool ok_sums = true;
  for (int i = 0; i < 6; ++i) ok_sums = ok_sums && is_finite(sums[i]);
  bool successful;
  double gmax;
  int iter;
  if (s.phase == 0) {
    if (!ok_sums) { s.done = CLC_TERM_FAILURE; return; }
    s.x_cost = s.initial_cost = sums[5];
    s.H[0] = sums[0]; s.H[1] = sums[1]; s.H[2] = sums[2];
    s.g[0] = sums[3]; s.g[1] = sums[4];
    s.scale[0] = 1.0 / (1.0 + sqrt(s.H[0]));
    s.scale[1] = 1.0 / (1.0 + sqrt(s.H[2]));
    successful = true;
    gmax = fmax(fabs(s.g[0]), fabs(s.g[1]));
    iter = 0;
  } else {
    const double cand_cost = ok_sums ? sums[5] : DBL_MAX;
    iter = s.iteration + 1;
    const double d0 = s.x[0] - s.cand[0], d1 = s.x[1] - s.cand[1];
    const double step_norm = sqrt(d0 * d0 + d1 * d1);
    const double cost_change = s.x_cost - cand_cost;
    if (step_norm <= ptol * (s.x_norm + ptol)) { s.done = CLC_TERM_CONVERGENCE_PARAMETER; return; }
    if (fabs(cost_change) <= ftol * s.x_cost) { s.done = CLC_TERM_CONVERGENCE_FUNCTION; return; }
    const double rel = cost_change / s.model_cost_change;
    gmax = 0.0;
    if (rel > min_rel) {
      s.x[0] = s.cand[0]; s.x[1] = s.cand[1];
      s.x_norm = sqrt(s.x[0] * s.x[0] + s.x[1] * s.x[1]);
      s.x_cost = cand_cost;
      s.H[0] = sums[0]; s.H[1] = sums[1]; s.H[2] = sums[2];
      s.g[0] = sums[3]; s.g[1] = sums[4];
      successful = true;
      gmax = fmax(fabs(s.g[0]), fabs(s.g[1]));
      const double q = 2.0 * rel - 1.0;
      double den = 1.0 - q * q * q;
      if (den < 1.0 / 3.0) den = 1.0 / 3.0;
      s.radius = s.radius / den;
      if (s.radius > 1e16) s.radius = 1e16;
      s.decrease_factor = 2.0;
      s.reuse_diagonal = 0;
    } else {
      successful = false;
      s.radius = s.radius / s.decrease_factor;
      s.decrease_factor *= 2.0;
      s.reuse_diagonal = 1;
    }
  }
  for (;;) {
    s.iteration = iter;
    if (iter >= max_num_iterations) { s.done = CLC_TERM_NO_CONVERGENCE; return; }
    if (successful && gmax <= gtol) { s.done = CLC_TERM_CONVERGENCE_GRADIENT; return; }
    if (!(s.radius > 1e-32)) { s.done = CLC_TERM_CONVERGENCE_MIN_RADIUS; return; }
    // Jacobi-scaled 2x2 system, (H_s + diag/radius) y = g_s, step = -y
    const double h00 = s.scale[0] * s.scale[0] * s.H[0], h01 = s.scale[0] * s.scale[1] * s.H[1],
                 h11 = s.scale[1] * s.scale[1] * s.H[2];
    const double g0 = s.scale[0] * s.g[0], g1 = s.scale[1] * s.g[1];
    if (!s.reuse_diagonal) {
      s.diag[0] = fmin(fmax(h00, min_diag), max_diag);
      s.diag[1] = fmin(fmax(h11, min_diag), max_diag);
    }
    const double a00 = h00 + s.diag[0] / s.radius, a11 = h11 + s.diag[1] / s.radius;
    // Cholesky of the 2x2
    bool ok = a00 > 0.0;
    const double l00 = sqrt(a00), l10 = h01 / l00, t = a11 - l10 * l10;
    ok = ok && (t > 0.0);
    const double l11 = sqrt(t);
    const double z0 = g0 / l00, z1 = (g1 - l10 * z0) / l11;
    const double y1 = z1 / l11, y0 = (z0 - l10 * y1) / l00;
    const double s0 = -y0, s1 = -y1;
    ok = ok && is_finite(s0) && is_finite(s1);
    s.reuse_diagonal = 1;
    double mcc = 0.0;
    if (ok) mcc = -(g0 * s0 + g1 * s1) - 0.5 * (s0 * (h00 * s0 + h01 * s1) + s1 * (h01 * s0 + h11 * s1));
    if (!(ok && mcc > 0.0)) {
      if (++s.num_invalid >= 5) { s.done = CLC_TERM_FAILURE; return; }
      s.radius = s.radius / s.decrease_factor;
      s.decrease_factor *= 2.0;
      s.reuse_diagonal = 1;
      successful = false;
      iter = s.iteration + 1;
      continue;
    }
    s.num_invalid = 0;
    s.cand[0] = s.x[0] + s0 * s.scale[0];
    s.cand[1] = s.x[1] + s1 * s.scale[1];
    s.model_cost_change = mcc;
    s.phase = 1;
    return;
  }
}

// ---- scan preparation: TranScanToPoints (reference src/utilities.cpp:181-215) + AutoGetLinePts (reference
//      src/selectScanPoints.cpp:17-190, without its OpenCV drawing) -------------------------------------------------------
// One beam of a LaserScan as the reference turns it into a point: (1000, 1000) marks an invalid range.
CLC_HD void scan_point(const float* ranges, int64_t i, double angle_min, double angle_increment, double range_min, double* x,
                       double* y) {
  const float range = ranges[i];
  if (range < 30.0 && range >= range_min) {
    const double ang = angle_min + (double)i * angle_increment;
    *x = (double)range * cos(ang);
    *y = (double)range * sin(ang);
  } else {
    *x = 1000.0;
    *y = 1000.0;
  }
}

// The longest continuous segment in the +-80 degree front sector; inclusive index range, or start = end = -1.
CLC_HD void auto_get_line_pts(const float* ranges, int64_t n, double angle_min, double angle_increment, double range_min,
                              int* seg_start_out, int* seg_end_out) {
  *seg_start_out = -1;
  *seg_end_out = -1;
  if (n <= 0) return;
  auto pt = [&](int64_t i, double* x, double* y) { scan_point(ranges, i, angle_min, angle_increment, range_min, x, y); };
  auto nrm = [&](int64_t i) {
    double x, y;
    pt(i, &x, &y);
    return sqrt(x * x + y * y);
  };
  const int64_t id = n / 2, delta = (int64_t)(80 / 0.3);
  const int64_t id_left = id + delta < n - 1 ? id + delta : n - 1;
  const int64_t id_right = id - delta > 0 ? id - delta : 0;
  const double dist_thre = 0.05, range_max = 100;
  const int skip = 3;
  int64_t best_cnt = -1;
  int64_t cur = id_right, next = cur + skip, seg_start = 0, seg_end = 0;
  bool new_seg = true;
  double d_cur = nrm(cur);
  for (int64_t i = id_right; i < id_left - skip; i += skip) {
    if (new_seg) { seg_start = cur; seg_end = next; new_seg = false; }
    const double d1 = d_cur, d2 = nrm(next);
    if (d1 < range_max && d2 < range_max) {
      if (fabs(d1 - d2) < dist_thre) {
        seg_end = next;
      } else {
        new_seg = true;
        double xs, ys, xe, ye;
        pt(seg_start, &xs, &ys);
        pt(seg_end, &xe, &ye);
        const double ds = sqrt(xs * xs + ys * ys), de = sqrt(xe * xe + ye * ye);
        if (sqrt((xs - xe) * (xs - xe) + (ys - ye) * (ys - ye)) > 0.2 && ds < 2 && de < 2 && seg_end - seg_start > 50) {
          int64_t s = seg_start, e = seg_end;
          for (int j = 1; j < 4; ++j) {
            const int64_t bp = seg_end + j;
            if (bp < n && fabs(de - nrm(bp)) < dist_thre) e = bp;
          }
          for (int j = -1; j > -4; --j) {
            const int64_t bp = seg_start + j;
            if (bp >= 0 && fabs(ds - nrm(bp)) < dist_thre) s = bp;
          }
          if (e - s > best_cnt) {
            best_cnt = e - s;
            *seg_start_out = (int)s;
            *seg_end_out = (int)e;
          }
        }
      }
      cur = next;
      d_cur = d2;
      next += skip;
    } else {
      if (d1 > range_max) {
        cur = next;
        d_cur = d2;
      }
      next += skip;
    }
  }
}

#if defined(__CUDACC__)
// one thread per scan: the walk is inherently sequential (about 180 steps), the batch is the parallelism
__global__ void clc_scan_segments_kernel(const float* __restrict__ ranges, int64_t n_scans, int64_t n_beams, double angle_min,
                                         double angle_increment, double range_min, int* __restrict__ seg_start,
                                         int* __restrict__ seg_end) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_scans) return;
  int a, b;
  auto_get_line_pts(ranges + s * n_beams, n_beams, angle_min, angle_increment, range_min, &a, &b);
  seg_start[s] = a;
  seg_end[s] = b;
}

// One warp fits one scan: points i in [b, e) at x[i * stride], y[i * stride] (stride 1: the SoA arrays of a problem; stride 3:
// the AoS xyz array of a std::vector<Eigen::Vector3d>).  line: start value in, fitted line out; info (optional, 4 doubles):
// termination, iterations, sweeps, final cost.
__device__ __forceinline__ void line_fit_warp(const double* __restrict__ x, const double* __restrict__ y, int stride, int64_t b,
                                              int64_t e, int max_num_iterations, double cauchy_a, double* __restrict__ line,
                                              double* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  Lm2 s;
  lm2_init(s, line[0], line[1]);
  const double inv_a2 = 1.0 / (cauchy_a * cauchy_a), a2 = cauchy_a * cauchy_a;
  while (!s.done) {
    const double m0 = s.cand[0], m1 = s.cand[1];
    double hxx = 0.0, hxy = 0.0, hyy = 0.0, gx = 0.0, gy = 0.0, prod = 1.0;
    int esum = 0;
    for (int64_t i = b + lane; i < e; i += 32) {
      const double px = x[i * stride], py = y[i * stride];
      const double r = fma(m0, px, fma(m1, py, 1.0));
      const double u = fma(r * inv_a2, r, 1.0);
      const double w = 1.0 / u;
      const double wx = w * px, wy = w * py, wr = w * r;
      hxx = fma(wx, px, hxx); hxy = fma(wx, py, hxy); hyy = fma(wy, py, hyy);
      gx = fma(wr, px, gx); gy = fma(wr, py, gy);
      prod *= u;
      const int hi = __double2hiint(prod);
      const int ex = (hi >> 20) - 1023;
      esum += ex;
      prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      hxx += __shfl_xor_sync(0xffffffffu, hxx, o);
      hxy += __shfl_xor_sync(0xffffffffu, hxy, o);
      hyy += __shfl_xor_sync(0xffffffffu, hyy, o);
      gx += __shfl_xor_sync(0xffffffffu, gx, o);
      gy += __shfl_xor_sync(0xffffffffu, gy, o);
      prod *= __shfl_xor_sync(0xffffffffu, prod, o);
      esum += __shfl_xor_sync(0xffffffffu, esum, o);
    }
    const double sums[6] = {hxx, hxy, hyy, gx, gy,
                            0.5 * a2 * (log(prod) + (double)esum * 0.693147180559945309417232121458)};
    lm2_update(s, sums, max_num_iterations);
  }
  if (lane == 0) {
    line[0] = s.x[0];
    line[1] = s.x[1];
    if (info != nullptr) {
      info[0] = (double)s.done;
      info[1] = (double)s.iteration;
      info[2] = (double)s.sweeps;
      info[3] = s.x_cost;
    }
  }
}

// batched: one warp per frame of a device-resident problem; lines[f*2..+2], info[f*4..+4]
__global__ void __launch_bounds__(256) clc_line_fit_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                          const int64_t* __restrict__ offsets, int64_t n_frames,
                                                          int max_num_iterations, double cauchy_a,
                                                          double* __restrict__ lines, double* __restrict__ info) {
  const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  line_fit_warp(x, y, 1, offsets[f], offsets[f + 1], max_num_iterations, cauchy_a, lines + 2 * f,
                info != nullptr ? info + 4 * f : nullptr);
}

// one scan straight from its AoS xyz array (the per-call shape of the reference's LineFittingCeres, main/calibr_offline.cpp:124):
// one warp, no problem object
__global__ void __launch_bounds__(32) clc_line_fit_single_kernel(const double* __restrict__ pts_xyz, int64_t n, int max_num_iterations,
                                                                double cauchy_a, double* __restrict__ line) {
  line_fit_warp(pts_xyz, pts_xyz + 1, 3, 0, n, max_num_iterations, cauchy_a, line, nullptr);
}
#endif

}  // namespace clc

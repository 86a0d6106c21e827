// clc_lm.cuh -- the Levenberg-Marquardt trust-region state machine, run ON THE DEVICE by one thread after every
// sweep of the fused residual+Jacobian+reduce kernel (either as the tail of that kernel's last block, or as its
// own single-thread kernel after the NCCL all-reduce when several ranks share the problem).
//
// It restates what ceres::Solve() does for the reference at src/LaseCamCalCeres.cpp:299-307 (DENSE_QR,
// max_num_iterations 100, every other option a Ceres default): Ceres' TrustRegionMinimizer loop with the
// LevenbergMarquardtStrategy, Jacobi scaling fixed at iteration 0, monotonic steps.  Two deliberate differences,
// both exact in exact arithmetic:
//   * the damped step solves the 6x6 normal equations (J_s^T J_s + D^2) y = J_s^T r by Cholesky instead of a
//     Householder QR of the (P+6)x6 matrix [J_s; D] -- the Jacobian is never materialised;
//   * every candidate point is evaluated with its Jacobian in the same sweep ("speculative" evaluation), so an
//     accepted step needs no second sweep: one LM iteration = one pass over the points.
#pragma once

#include "clc_math.cuh"
#include "../../include/clc_b200.h"

namespace clc {

constexpr int kTraceMax = 256;
constexpr int kNumSums = 28;  // 21 upper-tri H + 6 g + 1 cost

// Hot state of the minimiser (about 600 bytes): staged through shared memory around lm_update so that the single
// thread running it does not pay a global-memory round trip per field.
struct LmCore {
  int done;            // CLC_TERM_*; 0 while running
  int phase;           // 0: the pending sweep evaluates the start point; 1: it evaluates a candidate
  int iteration;       // index of the last finalised iteration
  int num_invalid;
  int reuse_diagonal;
  int n_trace;
  int num_successful;
  int num_unsuccessful;
  int sweeps;
  int pad0;
  double x[7];
  double cand[7];      // the pose the next sweep evaluates
  double x_cost, x_norm;
  double H[21], g[6];  // at x: loss-corrected, unscaled
  double scale[6], diag[6];
  double radius, decrease_factor, model_cost_change;
  double initial_cost;
  clc_lm_options opt;
};
static_assert(sizeof(LmCore) % 8 == 0, "LmCore is copied as 8-byte words");
constexpr int kLmCoreWords = (int)(sizeof(LmCore) / 8);

struct LmState {
  LmCore core;
  clc_lm_iteration trace[kTraceMax];
};

// -DCLC_LM_PROFILE: clock stamps at the section boundaries of the last lm_update that ran on the device (experiment builds only;
// read back with clc_debug_lm_profile)
#if defined(CLC_LM_PROFILE) && defined(__CUDACC__)
__device__ long long g_lm_profile[16];
#endif
#if defined(CLC_LM_PROFILE) && defined(__CUDA_ARCH__)
#define CLC_LM_STAMP(i) g_lm_profile[i] = clock64()
#else
#define CLC_LM_STAMP(i) ((void)0)
#endif

CLC_HD double norm7(const double* a) {
  double s = 0.0;
  for (int i = 0; i < 7; ++i) s += a[i] * a[i];
  return sqrt(s);
}

// Ceres EvaluateGradientAndJacobian: |x - Plus(x, -g)|_inf
CLC_HD double gradient_max_norm(const double* x, const double* g) {
  double ng[6], xp[7], m = 0.0;
  for (int i = 0; i < 6; ++i) ng[i] = -g[i];
  pose_plus(x, ng, xp);
  for (int i = 0; i < 7; ++i) {
    const double d = fabs(x[i] - xp[i]);
    if (d > m) m = d;
  }
  return m;
}

CLC_HD void lm_record(LmCore* s, clc_lm_iteration* trace, const clc_lm_iteration& it) {
  if (s->n_trace < kTraceMax) trace[s->n_trace] = it;
  s->n_trace++;
}

CLC_HD void lm_init(LmCore* s, const double* pose7, const clc_lm_options& opt) {
  s->done = 0; s->phase = 0; s->iteration = 0; s->num_invalid = 0; s->reuse_diagonal = 0; s->n_trace = 0;
  s->num_successful = 0; s->num_unsuccessful = 0; s->sweeps = 0; s->pad0 = 0;
  for (int i = 0; i < 7; ++i) { s->x[i] = pose7[i]; s->cand[i] = pose7[i]; }
  s->x_cost = 0.0;
  s->x_norm = norm7(pose7);
  s->radius = opt.initial_trust_region_radius;
  s->decrease_factor = 2.0;
  s->model_cost_change = 0.0;
  s->initial_cost = 0.0;
  s->opt = opt;
}

// Consumes the 28 sums of the sweep that has just evaluated s->cand and advances the minimiser until it either
// terminates (s->done != 0) or has a new candidate in s->cand for the next sweep.
CLC_HD void lm_update(LmCore* s, clc_lm_iteration* trace, const double* sums) {
  if (s->done) return;
  CLC_LM_STAMP(0);
  s->sweeps++;
  const clc_lm_options& o = s->opt;
  clc_lm_iteration last;
  last.reserved = 0;
  // Ceres rejects an evaluation that produced a non-finite residual or Jacobian entry (residual_block.cc
  // IsArrayValid): at the start point that is a FAILURE, at a candidate it is "a step with infinite cost".
  bool sums_ok = true;
  for (int i = 0; i < kNumSums; ++i) sums_ok = sums_ok && is_finite(sums[i]);
  if (s->phase == 0) {
    // ---- iteration 0 (Ceres: IterationZero) ----
    if (!sums_ok) { s->done = CLC_TERM_FAILURE; return; }
    s->x_cost = sums[27];
    for (int i = 0; i < 21; ++i) s->H[i] = sums[i];
    for (int i = 0; i < 6; ++i) s->g[i] = sums[21 + i];
    for (int k = 0; k < 6; ++k) s->scale[k] = o.jacobi_scaling ? 1.0 / (1.0 + sqrt(s->H[tri(k, k)])) : 1.0;
    s->initial_cost = s->x_cost;
    last.iteration = 0; last.step_is_valid = 1; last.step_is_successful = 1;
    last.cost = s->x_cost; last.cost_change = 0.0; last.gradient_max_norm = gradient_max_norm(s->x, s->g);
    last.step_norm = 0.0; last.relative_decrease = 0.0; last.trust_region_radius = s->radius;
  } else {
    // ---- a candidate has been evaluated ----
    const double cand_cost = sums_ok ? sums[27] : DBL_MAX;
    last.iteration = s->iteration + 1; last.step_is_valid = 1; last.step_is_successful = 0;
    double d[7];
    for (int i = 0; i < 7; ++i) d[i] = s->x[i] - s->cand[i];
    last.step_norm = norm7(d);
    last.cost_change = s->x_cost - cand_cost;
    last.cost = cand_cost;
    last.gradient_max_norm = 0.0; last.relative_decrease = 0.0; last.trust_region_radius = s->radius;
    // Ceres: ParameterToleranceReached
    if (last.step_norm <= o.parameter_tolerance * (s->x_norm + o.parameter_tolerance)) {
      s->done = CLC_TERM_CONVERGENCE_PARAMETER;
      lm_record(s, trace, last);
      return;
    }
    // Ceres: FunctionToleranceReached (tested before the accept/reject decision; the candidate is not applied)
    if (fabs(last.cost_change) <= o.function_tolerance * s->x_cost) {
      s->done = CLC_TERM_CONVERGENCE_FUNCTION;
      lm_record(s, trace, last);
      return;
    }
    last.relative_decrease = last.cost_change / s->model_cost_change;
    if (last.relative_decrease > o.min_relative_decrease) {
      // Ceres: HandleSuccessfulStep + LevenbergMarquardtStrategy::StepAccepted
      for (int i = 0; i < 7; ++i) s->x[i] = s->cand[i];
      s->x_norm = norm7(s->x);
      s->x_cost = cand_cost;
      for (int i = 0; i < 21; ++i) s->H[i] = sums[i];
      for (int i = 0; i < 6; ++i) s->g[i] = sums[21 + i];
      last.step_is_successful = 1;
      last.gradient_max_norm = gradient_max_norm(s->x, s->g);
      const double q = 2.0 * last.relative_decrease - 1.0;
      double den = 1.0 - q * q * q;
      if (den < 1.0 / 3.0) den = 1.0 / 3.0;
      s->radius = s->radius / den;
      if (s->radius > o.max_trust_region_radius) s->radius = o.max_trust_region_radius;
      s->decrease_factor = 2.0;
      s->reuse_diagonal = 0;
    } else {
      // Ceres: HandleUnsuccessfulStep + StepRejected
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->reuse_diagonal = 1;
    }
  }

  CLC_LM_STAMP(1);  // accept / reject decided (incl. gradient_max_norm of an accepted step)
  for (;;) {
    // ---- Ceres: FinalizeIterationAndCheckIfMinimizerCanContinue ----
    if (last.step_is_successful) s->num_successful++; else s->num_unsuccessful++;
    last.trust_region_radius = s->radius;
    lm_record(s, trace, last);
    s->iteration = last.iteration;
    if (last.iteration >= o.max_num_iterations) { s->done = CLC_TERM_NO_CONVERGENCE; return; }
    if (last.step_is_successful && last.gradient_max_norm <= o.gradient_tolerance) {
      s->done = CLC_TERM_CONVERGENCE_GRADIENT;
      return;
    }
    if (!(s->radius > o.min_trust_region_radius)) { s->done = CLC_TERM_CONVERGENCE_MIN_RADIUS; return; }

    CLC_LM_STAMP(2);  // iteration recorded, termination tests done
    // ---- Ceres: LevenbergMarquardtStrategy::ComputeStep on the Jacobi-scaled system ----
    double Hs[36], gs[6], A[36], step[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      gs[i] = s->scale[i] * s->g[i];
#pragma unroll
      for (int j = i; j < 6; ++j) {
        const double v = s->scale[i] * s->scale[j] * s->H[tri(i, j)];
        Hs[i * 6 + j] = v;
        Hs[j * 6 + i] = v;
      }
    }
    if (!s->reuse_diagonal)
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        double dd = Hs[k * 6 + k];
        dd = dd > o.min_lm_diagonal ? dd : o.min_lm_diagonal;
        dd = dd < o.max_lm_diagonal ? dd : o.max_lm_diagonal;
        s->diag[k] = dd;
      }
#pragma unroll
    for (int i = 0; i < 36; ++i) A[i] = Hs[i];
    const double inv_radius = 1.0 / s->radius;
#pragma unroll
    for (int k = 0; k < 6; ++k) A[k * 6 + k] += s->diag[k] * inv_radius;  // D^2 = diag / radius
    CLC_LM_STAMP(3);  // scaled, damped system built
    bool ok = chol6_solve(A, gs, step);
    CLC_LM_STAMP(4);  // Cholesky solve done
    s->reuse_diagonal = 1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      if (!is_finite(step[k])) ok = false;
      step[k] = -step[k];
    }
    // Ceres: model_cost_change = -(J s)^T (r + J s / 2) = -g_s.s - 1/2 s^T H_s s
    double mcc = 0.0;
    if (ok) {
      double gs_s = 0.0, sHs = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        gs_s += gs[i] * step[i];
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) r += Hs[i * 6 + j] * step[j];
        sHs += step[i] * r;
      }
      mcc = -gs_s - 0.5 * sHs;
    }
    if (!(ok && mcc > 0.0)) {
      // ---- Ceres: HandleInvalidStep ----
      if (++s->num_invalid >= o.max_num_consecutive_invalid_steps) { s->done = CLC_TERM_FAILURE; return; }
      s->radius = s->radius / s->decrease_factor;
      s->decrease_factor *= 2.0;
      s->reuse_diagonal = 1;
      const double prev_gmax = last.gradient_max_norm;
      last.iteration = s->iteration + 1; last.step_is_valid = 0; last.step_is_successful = 0;
      last.cost = s->x_cost; last.cost_change = 0.0; last.gradient_max_norm = prev_gmax;
      last.step_norm = 0.0; last.relative_decrease = 0.0;
      continue;
    }
    s->num_invalid = 0;
    double delta[6];
    for (int k = 0; k < 6; ++k) delta[k] = step[k] * s->scale[k];
    CLC_LM_STAMP(5);  // model cost change done
    pose_plus(s->x, delta, s->cand);
    s->model_cost_change = mcc;
    s->phase = 1;
    CLC_LM_STAMP(6);  // candidate pose done
    return;
  }
}

}  // namespace clc

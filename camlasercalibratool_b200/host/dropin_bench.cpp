// dropin_bench.cpp -- end-to-end timing THROUGH THE REFERENCE'S OWN SIGNATURE.
//
// What a user of MegviiRobot/CamLaserCalibraTool calls is
//     CamLaserCalibration(std::vector<Oberserve> obs, Eigen::Matrix4d& Tcl, false)
// (reference main/calibr_simulation.cpp:130, main/calibr_offline.cpp:170; declared include/LaseCamCalCeres.h:28), with
// every frame's points in its own pageable std::vector<Eigen::Vector3d>.  This driver builds exactly that object at the
// benchmark sizes (boards from the library's device generator, read back into std::vector<Oberserve>), calls the drop-in
// (host/LaseCamCalB200.cpp) and times the call with a host clock: marshal + gather/pack + PCIe + HBM layout + the LM
// solve + the analysis tail + tear-down are all inside.  bench.py reports the result as `e2e`.
//
//   clc_dropin_bench <frames> <beams> <sigma> <seed> <steps> <warmup> [edges]
// Devices: CLC_DEVICES (see include/clc_b200.h).  Output: one line "CLC_DROPIN_JSON {...}".
//
// Two call shapes are timed:
//   moved   CamLaserCalibration(std::move(copy), Tcl, false)   the by-value parameter is move-constructed
//   lvalue  CamLaserCalibration(obs, Tcl, false)               what the reference's callers write: the by-value signature
//                                                              makes the COMPILER deep-copy all points (one malloc + memcpy
//                                                              per frame and point set) before the callee runs
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <unistd.h>

#include "LaseCamCalCeres.h"
#include "clc_b200.h"

extern "C" void clc_dropin_last_phases(double out[7]);  // host/LaseCamCalB200.cpp

namespace {

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

double median(std::vector<double> v) {
  std::sort(v.begin(), v.end());
  return v.empty() ? 0.0 : v[v.size() / 2];
}

// stdout of the drop-in (its Ceres-style report) is not what is being measured: park it on /dev/null during the calls
struct QuietStdout {
  int saved = -1;
  QuietStdout() {
    std::fflush(stdout);
    saved = dup(1);
    FILE* f = std::fopen("/dev/null", "w");
    if (f) { dup2(fileno(f), 1); std::fclose(f); }
  }
  ~QuietStdout() {
    std::fflush(stdout);
    if (saved >= 0) { dup2(saved, 1); close(saved); }
  }
};

}  // namespace

int main(int argc, char** argv) {
  const int64_t frames = argc > 1 ? std::atoll(argv[1]) : 10000;
  const int64_t beams = argc > 2 ? std::atoll(argv[2]) : 1000;
  const double sigma = argc > 3 ? std::atof(argv[3]) : 0.01;
  const uint64_t seed = argc > 4 ? (uint64_t)std::atoll(argv[4]) : 7;
  const int steps = argc > 5 ? std::atoi(argv[5]) : 10;
  const int warmup = argc > 6 ? std::atoi(argv[6]) : 3;
  const int edges = argc > 7 ? std::atoi(argv[7]) : 0;

  int devices[16], n_devices = 0;
  if (clc_default_devices(devices, 16, &n_devices) != CLC_OK) {
    std::fprintf(stderr, "devices: %s\n", clc_last_error());
    return 2;
  }

  // ---- the observations, as the reference's callers hold them ----
  std::vector<Oberserve> obs((size_t)frames);
  {
    clc_synthetic_desc d;
    std::memset(&d, 0, sizeof(d));
    d.n_frames_total = frames; d.frame_begin = 0; d.frame_end = frames; d.beams = beams; d.seed = seed; d.sigma = sigma;
    d.with_edges = edges; d.use_loss = 1; d.cauchy_a = 0.05; d.device = devices[0];
    clc_problem* gen = nullptr;
    if (clc_problem_create_synthetic(&gen, &d) != CLC_OK) {
      std::fprintf(stderr, "generator: %s\n", clc_last_error());
      return 2;
    }
    std::vector<double> fp(7 * (size_t)frames), pts(3 * (size_t)(frames * beams)), ep(6 * (size_t)frames);
    std::vector<int64_t> off((size_t)frames + 1);
    if (clc_problem_download(gen, fp.data(), off.data(), pts.data(), edges ? ep.data() : nullptr, nullptr) != CLC_OK) {
      std::fprintf(stderr, "download: %s\n", clc_last_error());
      return 2;
    }
    clc_problem_destroy(gen);
    for (int64_t f = 0; f < frames; ++f) {
      Oberserve& ob = obs[(size_t)f];
      ob.tagPose_Qca = Eigen::Quaterniond(fp[7 * f + 3], fp[7 * f], fp[7 * f + 1], fp[7 * f + 2]);  // (w, x, y, z)
      ob.tagPose_tca = Eigen::Vector3d(fp[7 * f + 4], fp[7 * f + 5], fp[7 * f + 6]);
      ob.points.reserve((size_t)(off[f + 1] - off[f]));
      for (int64_t j = off[f]; j < off[f + 1]; ++j) ob.points.push_back(Eigen::Vector3d(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]));
      ob.points_on_line = ob.points;  // as reference main/calibr_simulation.cpp:101-102
      if (edges && !ob.points.empty()) {
        ob.points.front() = Eigen::Vector3d(ep[6 * f], ep[6 * f + 1], ep[6 * f + 2]);
        ob.points.back() = Eigen::Vector3d(ep[6 * f + 3], ep[6 * f + 4], ep[6 * f + 5]);
      }
    }
  }
  const int64_t n_points = frames * beams;
  const bool lfd = edges != 0;  // the edge residuals exist only with use_linefitting_data && use_boundary_constraint

  // ---- what one call computes: the same solve through the C ABI (sweep count, reference result) ----
  int lm_sweeps = 0, lm_iterations = 0, termination = 0;
  double ref_pose[7] = {0, 0, 0, 0, 0, 0, 1}, lm_device_ms = 0.0;
  int64_t h2d_bytes = 0;
  int pack_threads = 0, upload_chunks = 0, upload_direct = 0;
  double upload_ms = 0.0, upload_pack_wait_ms = 0.0;  // (first upload of the process: includes the one-time pinned-ring allocation)
  {
    std::vector<double> fp(7 * (size_t)frames), ep;
    std::vector<const double*> fpts((size_t)frames);
    std::vector<int64_t> cnt((size_t)frames);
    if (edges) ep.resize(6 * (size_t)frames);
    for (int64_t f = 0; f < frames; ++f) {
      const Oberserve& ob = obs[(size_t)f];
      fp[7 * f] = ob.tagPose_Qca.x(); fp[7 * f + 1] = ob.tagPose_Qca.y(); fp[7 * f + 2] = ob.tagPose_Qca.z(); fp[7 * f + 3] = ob.tagPose_Qca.w();
      fp[7 * f + 4] = ob.tagPose_tca.x(); fp[7 * f + 5] = ob.tagPose_tca.y(); fp[7 * f + 6] = ob.tagPose_tca.z();
      const std::vector<Eigen::Vector3d>& p = lfd ? ob.points_on_line : ob.points;
      fpts[(size_t)f] = reinterpret_cast<const double*>(p.data());
      cnt[(size_t)f] = (int64_t)p.size();
      if (edges) {
        const Eigen::Vector3d &a = ob.points.front(), &b = ob.points.back();
        double* e = &ep[6 * (size_t)f];
        e[0] = a.x(); e[1] = a.y(); e[2] = a.z(); e[3] = b.x(); e[4] = b.y(); e[5] = b.z();
      }
    }
    clc_gather_desc gd = {frames, fp.data(), fpts.data(), cnt.data(), edges ? ep.data() : nullptr, 1, 0.05, -1};
    clc_group* g = nullptr;
    if (clc_group_create_gather(&g, &gd, devices, n_devices) != CLC_OK) {
      std::fprintf(stderr, "group: %s\n", clc_last_error());
      return 2;
    }
    clc_upload_last_stats(&upload_ms, &upload_pack_wait_ms, &h2d_bytes, &upload_chunks, &pack_threads, &upload_direct);
    clc_lm_summary s;
    if (clc_group_solve_lm(g, ref_pose, nullptr, &s, nullptr, 0) != CLC_OK) {
      std::fprintf(stderr, "solve: %s\n", clc_last_error());
      return 2;
    }
    lm_sweeps = s.num_sweeps; lm_iterations = s.num_iterations - 1; termination = s.termination; lm_device_ms = s.device_ms;
    clc_group_destroy(g);
  }
  double ref_T[16];
  clc_pose7_to_T(ref_pose, ref_T);

  // ---- raw PCIe time of the bytes one call uploads (pinned -> device, same size), the yardstick for the upload ----
  float h2d_ms_raw[8] = {0};
  const int h2d_reps = 5;
  double raw_h2d_ms = 0.0;
  if (clc_bench_h2d(h2d_bytes / std::max(1, n_devices), devices[0], h2d_reps, h2d_ms_raw) == CLC_OK) {
    std::vector<double> v(h2d_ms_raw, h2d_ms_raw + h2d_reps);
    raw_h2d_ms = median(v);
  }

  // ---- what the by-value signature costs the CALLER, with nothing of ours involved: deep copy and destruction ----
  double copy_ms = 0.0, destroy_ms = 0.0;
  {
    std::vector<double> c, d;
    for (int it = 0; it < 3; ++it) {
      const double t0 = now_ms();
      std::vector<Oberserve>* tmp = new std::vector<Oberserve>(obs);
      const double t1 = now_ms();
      delete tmp;
      const double t2 = now_ms();
      c.push_back(t1 - t0);
      d.push_back(t2 - t1);
    }
    copy_ms = median(c);
    destroy_ms = median(d);
  }

  // ---- timed calls ----
  // body_*: entry to return of CamLaserCalibration() (measured inside the drop-in: marshal + gather/pack + PCIe + HBM layout +
  // LM solve + report + analysis tail + tear-down) -- the headline.  call_*: the whole call expression on the caller's clock,
  // which adds what the by-value `obs` parameter costs the caller (move or deep copy before entry, destruction after return).
  std::vector<double> body, call_moved, call_lvalue, body_lvalue;
  std::vector<double> phase[7];
  double max_dev = 0.0;
  for (int it = 0; it < warmup + steps; ++it) {
    Eigen::Matrix4d Tcl = Eigen::Matrix4d::Identity();
    std::vector<Oberserve> copy = obs;  // outside the timed region
    double t0, t1;
    {
      QuietStdout quiet;
      t0 = now_ms();
      CamLaserCalibration(std::move(copy), Tcl, lfd, edges != 0);
      t1 = now_ms();
    }
    double ph[7];
    clc_dropin_last_phases(ph);
    if (it >= warmup) {
      call_moved.push_back(t1 - t0);
      body.push_back(ph[6]);
      for (int k = 0; k < 7; ++k) phase[k].push_back(ph[k]);
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) max_dev = std::max(max_dev, std::fabs(Tcl(r, c) - ref_T[r * 4 + c]));
  }
  const int lv_steps = std::max(1, std::min(steps, 3));
  for (int it = 0; it < 1 + lv_steps; ++it) {
    Eigen::Matrix4d Tcl = Eigen::Matrix4d::Identity();
    double t0, t1;
    {
      QuietStdout quiet;
      t0 = now_ms();
      CamLaserCalibration(obs, Tcl, lfd, edges != 0);
      t1 = now_ms();
    }
    double ph[7];
    clc_dropin_last_phases(ph);
    if (it >= 1) { call_lvalue.push_back(t1 - t0); body_lvalue.push_back(ph[6]); }
  }
  double sum = 0.0;
  for (double v : body) sum += v;
  const double mean_body = sum / (double)body.size();
  double phase_sum[7];  // per-phase medians over the timed calls
  for (int k = 0; k < 7; ++k) phase_sum[k] = median(phase[k]);
  const int sweeps_per_call = lm_sweeps + 1;  // + the un-robustified information sweep of the analysis tail (:318-362)
  std::printf(
      "CLC_DROPIN_JSON {\"frames\": %lld, \"beams\": %lld, \"points\": %lld, \"edges\": %d, \"n_devices\": %d, \"steps\": %d, \"warmup\": %d, "
      "\"body_ms_mean\": %.6f, \"body_ms_median\": %.6f, \"body_ms_min\": %.6f, \"body_ms_max\": %.6f, "
      "\"phases_ms_median\": {\"marshal\": %.4f, \"upload\": %.4f, \"solve\": %.4f, \"report\": %.4f, \"information\": %.4f, \"destroy\": %.4f}, "
      "\"call_expr_moved_ms_median\": %.6f, \"call_expr_lvalue_ms_median\": %.6f, \"body_ms_in_lvalue_calls_median\": %.6f, "
      "\"caller_copy_of_obs_ms\": %.6f, \"caller_destruction_of_obs_ms\": %.6f, "
      "\"sweeps_per_call\": %d, \"lm_iterations\": %d, \"termination\": %d, "
      "\"lm_device_ms\": %.6f, \"h2d_bytes_per_call\": %lld, \"d2h_bytes_per_call\": %d, \"raw_h2d_ms_same_bytes\": %.6f, "
      "\"upload_chunks\": %d, \"pack_threads\": %d, \"upload_direct\": %d, \"max_abs_dev_vs_c_abi_solve\": %.3e}\n",
      (long long)frames, (long long)beams, (long long)n_points, edges, n_devices, steps, warmup, mean_body, median(body),
      *std::min_element(body.begin(), body.end()), *std::max_element(body.begin(), body.end()),
      phase_sum[0], phase_sum[1], phase_sum[2], phase_sum[3], phase_sum[4], phase_sum[5],
      median(call_moved), median(call_lvalue), median(body_lvalue), copy_ms, destroy_ms,
      sweeps_per_call, lm_iterations, termination, lm_device_ms, (long long)h2d_bytes,
      (int)clc_solve_readback_bytes() + 28 * 8, raw_h2d_ms, upload_chunks, pack_threads, upload_direct, max_dev);
  return max_dev < 1e-9 ? 0 : 3;
}

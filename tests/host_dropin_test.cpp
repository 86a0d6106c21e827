// Test driver for the C++ drop-in (camlasercalibratool_b200/host/LaseCamCalB200.cpp): does what the reference's
// callers do -- main/calibr_simulation.cpp:116-130 (identity start, LM, use_linefitting_data = false) and
// main/calibr_offline.cpp:166-170 (closed form, invert, LM) -- through the reference's own function signatures.
// Observations come from the library's device generator, read back into std::vector<Oberserve>.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "LaseCamCalCeres.h"
#include "clc_b200.h"

static void invert_rigid(const Eigen::Matrix4d& T, Eigen::Matrix4d& out) {
  out.setIdentity();
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) out(r, c) = T(c, r);
    out(r, 3) = -(T(0, r) * T(0, 3) + T(1, r) * T(1, 3) + T(2, r) * T(2, 3));
  }
}

static void dump(const char* tag, const Eigen::Matrix4d& T) {
  std::printf("%s", tag);
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) std::printf(" %.17g", T(r, c));
  std::printf("\n");
}

int main(int argc, char** argv) {
  const int64_t frames = argc > 1 ? std::atoll(argv[1]) : 50;
  const int64_t beams = argc > 2 ? std::atoll(argv[2]) : 180;
  const double sigma = argc > 3 ? std::atof(argv[3]) : 0.0;
  const int edges = argc > 4 ? std::atoi(argv[4]) : 0;
  clc_synthetic_desc d = {frames, 0, frames, beams, 1, sigma, edges, 1, 0.05, -1};
  clc_problem* gen = nullptr;
  if (clc_problem_create_synthetic(&gen, &d) != CLC_OK) {
    std::fprintf(stderr, "generator: %s\n", clc_last_error());
    return 2;
  }
  std::vector<double> fp(7 * frames), pts(3 * frames * beams), ep(6 * frames);
  std::vector<int64_t> off(frames + 1);
  clc_problem_download(gen, fp.data(), off.data(), pts.data(), edges ? ep.data() : nullptr, nullptr);
  clc_problem_destroy(gen);

  std::vector<Oberserve> obs(frames);
  for (int64_t f = 0; f < frames; ++f) {
    Oberserve& ob = obs[f];
    ob.tagPose_Qca = Eigen::Quaterniond(fp[7 * f + 3], fp[7 * f], fp[7 * f + 1], fp[7 * f + 2]);  // (w, x, y, z)
    ob.tagPose_tca = Eigen::Vector3d(fp[7 * f + 4], fp[7 * f + 5], fp[7 * f + 6]);
    for (int64_t j = off[f]; j < off[f + 1]; ++j) ob.points.push_back(Eigen::Vector3d(pts[3 * j], pts[3 * j + 1], pts[3 * j + 2]));
    ob.points_on_line = ob.points;  // as reference main/calibr_simulation.cpp:101-102
    if (edges) {  // put the edge points where the reference takes them from: points.front() / points.back()
      ob.points.front() = Eigen::Vector3d(ep[6 * f], ep[6 * f + 1], ep[6 * f + 2]);
      ob.points.back() = Eigen::Vector3d(ep[6 * f + 3], ep[6 * f + 4], ep[6 * f + 5]);
    }
  }

  // calibr_simulation.cpp: Tcl = identity^-1, CamLaserCalibration(obs, Tcl, false)
  Eigen::Matrix4d Tcl_sim = Eigen::Matrix4d::Identity();
  CamLaserCalibration(obs, Tcl_sim, false);
  dump("RESULT_SIM_TCL", Tcl_sim);

  // calibr_offline.cpp: closed form -> invert -> LM
  Eigen::Matrix4d Tlc0 = Eigen::Matrix4d::Identity();
  CamLaserCalClosedSolution(obs, Tlc0);
  dump("RESULT_CLOSED_TLC", Tlc0);
  Eigen::Matrix4d Tcl_off;
  invert_rigid(Tlc0, Tcl_off);
  CamLaserCalibration(obs, Tcl_off, false);
  dump("RESULT_OFFLINE_TCL", Tcl_off);

  if (edges) {  // the "hidden feature": both flags set
    Eigen::Matrix4d Tcl_e;
    invert_rigid(Tlc0, Tcl_e);
    CamLaserCalibration(obs, Tcl_e, true, true);
    dump("RESULT_EDGES_TCL", Tcl_e);
  }
  return 0;
}

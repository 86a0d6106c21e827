// Test-only harness: compiles the product's host/device (CLC_HD) headers with g++ so that the exact source the GPU
// runs -- the LM state machine, the moment expansion, the plane and generator math -- can be checked against the
// oracle on a machine without a GPU.  Never shipped, never linked into libclc_b200.so.
#include <cstring>
#include <vector>

#include "../camlasercalibratool_b200/csrc/clc_expand.cuh"
#include "../camlasercalibratool_b200/csrc/clc_camera.cuh"
#include "../camlasercalibratool_b200/csrc/clc_linefit.cuh"
#include "../camlasercalibratool_b200/csrc/clc_lm.cuh"

extern "C" {

int harness_lm_state_size() { return (int)sizeof(clc::LmState); }
void harness_lm_init(void* st, const double* pose7, const clc_lm_options* opt) {
  clc::lm_init(&static_cast<clc::LmState*>(st)->core, pose7, *opt);
}
void harness_lm_update(void* st, const double* sums28) {
  clc::LmState* s = static_cast<clc::LmState*>(st);
  clc::lm_update(&s->core, s->trace, sums28);
}
int harness_lm_done(const void* st) { return static_cast<const clc::LmState*>(st)->core.done; }
int harness_lm_ntrace(const void* st) { return static_cast<const clc::LmState*>(st)->core.n_trace; }
void harness_lm_cand(const void* st, double* out) { std::memcpy(out, static_cast<const clc::LmState*>(st)->core.cand, 56); }
void harness_lm_x(const void* st, double* out) { std::memcpy(out, static_cast<const clc::LmState*>(st)->core.x, 56); }
void harness_lm_trace(const void* st, int i, clc_lm_iteration* out) { *out = static_cast<const clc::LmState*>(st)->trace[i]; }
int harness_lm_sweeps(const void* st) { return static_cast<const clc::LmState*>(st)->core.sweeps; }

// moments of one piece (computed by the caller) -> the 28 sums, through the same code path as the kernel
void harness_expand_lm(const double* plane, const double* pose7, double count, const double* S10, int use_loss,
                       double cost_term, double a2, double* out28) {
  clc::PoseConsts pc;
  clc::make_pose_consts(pose7, &pc);
  double m[3], c;
  clc::frame_consts(pc, plane, m, &c);
  clc::expand_lm(plane, m, c, 1.0 / count, S10, use_loss != 0, cost_term, a2, out28);
}
// one residual added directly to the 28 sums: the per-residual code of the one-cluster kernel (csrc/clc_small.cuh)
void harness_accumulate_residual(const double* plane, const double* pose7, const double* xyz, double count, int use_loss,
                                 double a2, double* acc28) {
  clc::PoseConsts pc;
  clc::make_pose_consts(pose7, &pc);
  clc::accumulate_residual(pc, plane, xyz[0], xyz[1], xyz[2], 1.0 / count, use_loss != 0, a2, 1.0 / a2, acc28);
}
void harness_frame_consts(const double* plane, const double* pose7, double* m3, double* c) {
  clc::PoseConsts pc;
  clc::make_pose_consts(pose7, &pc);
  clc::frame_consts(pc, plane, m3, c);
}
void harness_expand_closed(const double* plane, const double* S10, double* out54) {
  clc::expand_closed_form(plane, S10, out54);
}
void harness_frame_plane(const double* fp, double* plane) { clc::frame_plane(fp, plane); }
void harness_edge_planes(const double* fp, double* p1, double* p2) { clc::edge_planes(fp, p1, p2); }
void harness_pose_plus(const double* x, const double* d, double* xp) { clc::pose_plus(x, d, xp); }
void harness_quat_to_rot(const double* q, double* R) { clc::quat_to_rot(q, R); }
void harness_rot_to_quat(const double* R, double* q) { clc::rot_to_quat(R, q); }
int harness_chol6_solve(const double* A, const double* b, double* y) { return clc::chol6_solve(A, b, y) ? 1 : 0; }
int harness_solve_linear(double* A, double* b, int n) { return clc::solve_linear(A, b, n) ? 1 : 0; }
double harness_equi_r(const double* k, double th) { return clc::equi_r(k, th); }
double harness_equi_theta_from_r(const double* k, double rn) { return clc::equi_theta_from_r(k, rn); }
void harness_philox(uint64_t seed, uint64_t lo, uint64_t hi, uint32_t* out) { clc::philox4x32(seed, lo, hi, out); }
void harness_gen_frame_pose(uint64_t seed, int64_t frame, int with_edges, double* fp) {
  clc::gen_frame_pose(seed, frame, with_edges != 0, fp);
}
int harness_gen_edge_points(const double* fp, double* ep) { return clc::gen_edge_points(fp, ep) ? 1 : 0; }
// exact-M points of one frame, as clc_gen_points_kernel computes them
void harness_gen_points(uint64_t seed, double sigma, int64_t frame, int64_t beams, const double* fp, double* pts) {
  double nl[3], dl, a = 0.0, b = 0.0;
  clc::gen_plane_laser(fp, nl, &dl);
  clc::gen_window(nl, dl, &a, &b);
  for (int64_t j = 0; j < beams; ++j) {
    const double theta = a + (b - a) * (((double)j + 0.5) / (double)beams);
    const double cx = cos(theta), sy = sin(theta);
    const double depth = -dl / (cx * nl[0] + sy * nl[1]) + clc::gen_noise(seed, sigma, frame, j);
    pts[3 * j] = depth * cx; pts[3 * j + 1] = depth * sy; pts[3 * j + 2] = 0.0;
  }
}

// ---- camera measurement chain (clc_camera.cuh) ----
static clc::CameraDesc make_cam(int model, const double* intr, double sigma, int rows, int cols, double tag, double spacing) {
  clc::CameraDesc c;
  c.model = model;
  for (int k = 0; k < 8; ++k) c.intr[k] = intr[k];
  c.pixel_sigma = sigma; c.grid_rows = rows; c.grid_cols = cols; c.tag_size = tag; c.tag_spacing = spacing;
  return c;
}
void harness_camera_project(int model, const double* intr, const double* P, double* uv) {
  clc::camera_project(make_cam(model, intr, 0, 6, 6, 0.055, 0.3), P, uv, uv + 1);
}
void harness_camera_lift(int model, const double* intr, const double* uv, double* xy) {
  clc::camera_lift_normalised(make_cam(model, intr, 0, 6, 6, 0.055, 0.3), uv[0], uv[1], xy, xy + 1);
}
void harness_grid_corners(int rows, int cols, double tag, double spacing, double* xy) {
  const double z[8] = {1, 1, 0, 0, 0, 0, 0, 0};
  const clc::CameraDesc c = make_cam(1, z, 0, rows, cols, tag, spacing);
  for (int i = 0; i < clc::grid_num_corners(c); ++i) clc::grid_corner(c, i, xy + 2 * i, xy + 2 * i + 1);
}
void harness_pixel_noise(uint64_t seed, double sigma, int64_t frame, int corner, double* n2) {
  clc::pixel_noise(seed, sigma, frame, corner, n2, n2 + 1);
}
int harness_pnp_planar(int n, const double* obj_xy, const double* img_uv, double* R9, double* t3) {
  auto obj = [&](int i, double* X, double* Y) { *X = obj_xy[2 * i]; *Y = obj_xy[2 * i + 1]; };
  auto img = [&](int i, double* u, double* v) { *u = img_uv[2 * i]; *v = img_uv[2 * i + 1]; };
  return clc::pnp_planar(n, obj, img, R9, t3) ? 1 : 0;
}
int harness_camera_estimate_pose(int model, const double* intr, double sigma, int rows, int cols, double tag, double spacing,
                                 uint64_t seed, int64_t frame, const double* fp_true, double* fp_est, float* uv_out) {
  return clc::camera_estimate_pose(make_cam(model, intr, sigma, rows, cols, tag, spacing), seed, frame, fp_true, fp_est, uv_out) ? 1 : 0;
}

int harness_gen_frame_pose_camera(int model, const double* intr, int rows, int cols, double tag, double spacing, int width,
                                  int height, uint64_t seed, int64_t frame, int with_edges, double* fp) {
  return clc::gen_frame_pose_camera(make_cam(model, intr, 0, rows, cols, tag, spacing), width, height, seed, frame,
                                    with_edges != 0, fp) ? 1 : 0;
}

int harness_estimate_pose_from_detections(int model, const double* intr, int rows, int cols, double tag, double spacing, int n_det,
                                          const int* ids, const float* corners, double* pose_wc) {
  std::vector<float> lifted(8 * (size_t)(n_det > 0 ? n_det : 1));
  return clc::estimate_pose_from_detections(make_cam(model, intr, 0, rows, cols, tag, spacing), n_det, ids, corners,
                                            lifted.data(), pose_wc) ? 1 : 0;
}

void harness_auto_get_line_pts(const float* ranges, int64_t n, double a0, double inc, double rmin, int* s, int* e) {
  clc::auto_get_line_pts(ranges, n, a0, inc, rmin, s, e);
}
// the 2-parameter state machine of the batched line fit, driven exactly as the kernel does
int harness_lm2_size() { return (int)sizeof(clc::Lm2); }
void harness_lm2_init(void* st, double m0, double m1) { clc::lm2_init(*static_cast<clc::Lm2*>(st), m0, m1); }
void harness_lm2_update(void* st, const double* sums6, int max_iter) { clc::lm2_update(*static_cast<clc::Lm2*>(st), sums6, max_iter); }
int harness_lm2_done(const void* st) { return static_cast<const clc::Lm2*>(st)->done; }
void harness_lm2_get(const void* st, double* cand2, double* x2, int* iteration, int* sweeps) {
  const clc::Lm2* s = static_cast<const clc::Lm2*>(st);
  cand2[0] = s->cand[0]; cand2[1] = s->cand[1]; x2[0] = s->x[0]; x2[1] = s->x[1];
  *iteration = s->iteration; *sweeps = s->sweeps;
}

}  // extern "C"

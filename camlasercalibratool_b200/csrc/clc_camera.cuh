// clc_camera.cuh -- the measurement chain that produces a board pose from an image in the reference, restated for
// the synthetic generator (gives BASELINE's "radtan pinhole" / "equi" config labels a meaning):
//
//   board corners (kalibr april grid, reference src/calcCamPose.cpp:48-139)
//     -> camera frame with the TRUE board pose
//     -> pixels: Camera::spaceToPlane  (pinhole + radial-tangential: camera_models/src/PinholeCamera.cc:428-451,554-572;
//                                       equidistant / Kannala-Brandt: camera_models/src/EquidistantCamera.cc:364-378,
//                                       camera_models/include/EquidistantCamera.h:153-161)
//     -> + Gaussian pixel noise (stands in for detection + cornerSubPix error)
//     -> normalised plane: Camera::liftProjective (PinholeCamera.cc:358-420: 8-step recursive undistortion;
//                                       EquidistantCamera.cc:342-357,632-734: smallest non-negative real root of the
//                                       odd polynomial theta + k2 theta^3 + ... = r)
//     -> float32 points, as the reference stores them in cv::Point2f / cv::Point3f (calcCamPose.cpp:284-292)
//     -> cv::solvePnP(K = I, no distortion) (calcCamPose.cpp:225): planar pose from a homography + Levenberg-Marquardt
//        refinement of the reprojection error
//   -> the ESTIMATED board pose (R_ca, t_ca) the calibration sees, while the laser hits the TRUE board.
//
// Everything is host/device code so that tests can run it on the CPU against the oracle, numpy and OpenCV itself.
#pragma once

#include "clc_math.cuh"

namespace clc {

enum CameraModel { kCameraNone = 0, kCameraPinholeRadtan = 1, kCameraEquidistant = 2 };

// intr[8]: pinhole  = fx, fy, cx, cy, k1, k2, p1, p2        (reference config/calibra_config_pinhole.yaml)
//          equidist = mu, mv, u0, v0, k2, k3, k4, k5        (reference config/calibra_config.yaml, KANNALA_BRANDT)
struct CameraDesc {
  int model;
  double intr[8];
  double pixel_sigma;   // std of the Gaussian corner noise, pixels
  int grid_rows, grid_cols;
  double tag_size, tag_spacing;  // metres, ratio (kalibr convention: pitch = tag_size * (1 + tag_spacing))
};

// ---- pinhole + radtan ----------------------------------------------------------------------------------------------
CLC_HD void radtan_distortion(const double* k, double x, double y, double* dx, double* dy) {  // PinholeCamera.cc:554-572
  const double k1 = k[4], k2 = k[5], p1 = k[6], p2 = k[7];
  const double mx2 = x * x, my2 = y * y, mxy = x * y, rho2 = mx2 + my2;
  const double rad = k1 * rho2 + k2 * rho2 * rho2;
  *dx = x * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2);
  *dy = y * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2);
}

CLC_HD bool radtan_no_distortion(const double* k) { return k[4] == 0.0 && k[5] == 0.0 && k[6] == 0.0 && k[7] == 0.0; }

// ---- equidistant ------------------------------------------------------------------------------------------------------
CLC_HD double equi_r(const double* k, double th) {  // EquidistantCamera.h:153-161 (k1 = 1)
  const double t2 = th * th;
  return th * (1.0 + t2 * (k[4] + t2 * (k[5] + t2 * (k[6] + t2 * k[7]))));
}

// smallest non-negative real root of r(theta) = rn (backprojectSymmetric): the reference takes it from the eigenvalues
// of the companion matrix; r is increasing from 0 on the branch that contains it, so a safeguarded Newton iteration
// started at 0 walks up to that same root.
CLC_HD double equi_theta_from_r(const double* k, double rn) {
  if (k[4] == 0.0 && k[5] == 0.0 && k[6] == 0.0 && k[7] == 0.0) return rn;
  double th = 0.0;
  for (int it = 0; it < 60; ++it) {
    const double t2 = th * th;
    const double f = equi_r(k, th) - rn;
    const double df = 1.0 + t2 * (3.0 * k[4] + t2 * (5.0 * k[5] + t2 * (7.0 * k[6] + t2 * 9.0 * k[7])));
    if (!(df > 1e-12)) break;  // flat / turning point: no root on this branch below it
    const double step = f / df;
    th -= step;
    if (th < 0.0) th = 0.0;
    if (fabs(step) < 1e-15 * (1.0 + fabs(th))) break;
  }
  return th;
}

// ---- Camera::spaceToPlane / Camera::liftProjective ------------------------------------------------------------------------
CLC_HD void camera_project(const CameraDesc& cam, const double* P, double* u, double* v) {
  const double* k = cam.intr;
  if (cam.model == kCameraEquidistant) {
    const double nrm = sqrt(P[0] * P[0] + P[1] * P[1] + P[2] * P[2]);
    const double theta = acos(P[2] / nrm), phi = atan2(P[1], P[0]);
    const double r = equi_r(k, theta);
    *u = k[0] * (r * cos(phi)) + k[2];
    *v = k[1] * (r * sin(phi)) + k[3];
  } else {
    double x = P[0] / P[2], y = P[1] / P[2];
    if (!radtan_no_distortion(k)) {
      double dx, dy;
      radtan_distortion(k, x, y, &dx, &dy);
      x += dx;
      y += dy;
    }
    *u = k[0] * x + k[2];
    *v = k[1] * y + k[3];
  }
}

// pixel -> point on the normalised image plane (x/z, y/z), as reference src/calcCamPose.cpp:284-292 uses the lifted ray
CLC_HD void camera_lift_normalised(const CameraDesc& cam, double u, double v, double* xn, double* yn) {
  const double* k = cam.intr;
  const double mx = (u - k[2]) / k[0], my = (v - k[3]) / k[1];  // m_inv_K11 * u + m_inv_K13
  if (cam.model == kCameraEquidistant) {
    const double rn = sqrt(mx * mx + my * my);
    const double phi = rn < 1e-10 ? 0.0 : atan2(my, mx);
    const double theta = equi_theta_from_r(k, rn);
    const double s = sin(theta), c = cos(theta);
    *xn = s * cos(phi) / c;
    *yn = s * sin(phi) / c;
  } else {
    double xu = mx, yu = my;
    if (!radtan_no_distortion(k)) {
      double dx, dy;
      radtan_distortion(k, mx, my, &dx, &dy);
      xu = mx - dx;
      yu = my - dy;
      for (int i = 1; i < 8; ++i) {  // "Recursive distortion model", n = 8
        radtan_distortion(k, xu, yu, &dx, &dy);
        xu = mx - dx;
        yu = my - dy;
      }
    }
    *xn = xu;
    *yn = yu;
  }
}

// ---- board corners: kalibr april grid, 4 corners per tag (reference src/calcCamPose.cpp:107-136) -------------------------
CLC_HD int grid_num_corners(const CameraDesc& cam) { return 4 * cam.grid_rows * cam.grid_cols; }
CLC_HD void grid_corner(const CameraDesc& cam, int idx, double* X, double* Y) {
  const int tag = idx >> 2, j = idx & 3;
  const int row = tag / cam.grid_cols, col = tag % cam.grid_cols;
  const double pitch = cam.tag_size * (1. + cam.tag_spacing);
  *X = pitch * col + ((j == 1 || j == 2) ? cam.tag_size : 0.0);
  *Y = pitch * row + ((j == 2 || j == 3) ? cam.tag_size : 0.0);
}

constexpr uint64_t kStreamPixel = (uint64_t)3 << 56;
CLC_HD void pixel_noise(uint64_t seed, double sigma, int64_t frame, int corner, double* nu, double* nv) {
  *nu = 0.0;
  *nv = 0.0;
  if (!(sigma > 0.0)) return;
  uint32_t o[4];
  philox4x32(seed, (uint64_t)frame, kStreamPixel | (uint64_t)corner, o);
  const double u1 = u53(o[0], o[1]), u2 = u53(o[2], o[3]);
  const double rad = sigma * sqrt(-2.0 * log(1.0 - u1));  // Box-Muller, both variates
  *nu = rad * cos(2.0 * kPi * u2);
  *nv = rad * sin(2.0 * kPi * u2);
}

// ---- planar PnP: what cv::solvePnP(SOLVEPNP_ITERATIVE) does for coplanar points with K = I --------------------------------
// obj: (X, Y) on the board plane (Z = 0), img: normalised image points; n >= 4.  R row-major, t.  false if degenerate.
// 1. homography by normalised DLT (8 unknowns, h33 = 1), 2. R, t from H = lambda [r1 r2 t], 3. Levenberg-Marquardt on the
// reprojection error with a left-multiplicative rotation update, run to convergence (OpenCV stops after <= 20 iterations
// at FLT_EPSILON; both end at the same minimum).
CLC_HD bool solve_linear(double* A, double* b, int n) {  // Gaussian elimination with partial pivoting, in place; x in b
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(A[r * n + c]) > fabs(A[piv * n + c])) piv = r;
    if (!(fabs(A[piv * n + c]) > 1e-300)) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) { const double t = A[c * n + k]; A[c * n + k] = A[piv * n + k]; A[piv * n + k] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * n + c] * inv;
      if (f == 0.0) continue;
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * b[k];
    b[r] = s / A[r * n + r];
  }
  return true;
}

template <class ObjFn, class ImgFn>
CLC_HD bool pnp_planar(int n, ObjFn obj, ImgFn img, double* R, double* t) {
  if (n < 4) return false;
  // ---- 1. Hartley-normalised DLT ----
  double mo[2] = {0, 0}, mi[2] = {0, 0};
  for (int i = 0; i < n; ++i) {
    double X, Y, u, v;
    obj(i, &X, &Y); img(i, &u, &v);
    mo[0] += X; mo[1] += Y; mi[0] += u; mi[1] += v;
  }
  for (int k = 0; k < 2; ++k) { mo[k] /= n; mi[k] /= n; }
  double so = 0.0, si = 0.0;
  for (int i = 0; i < n; ++i) {
    double X, Y, u, v;
    obj(i, &X, &Y); img(i, &u, &v);
    so += sqrt((X - mo[0]) * (X - mo[0]) + (Y - mo[1]) * (Y - mo[1]));
    si += sqrt((u - mi[0]) * (u - mi[0]) + (v - mi[1]) * (v - mi[1]));
  }
  if (!(so > 0.0) || !(si > 0.0)) return false;
  so = 1.41421356237309515 * n / so;
  si = 1.41421356237309515 * n / si;
  double A[64], b[8];
  for (int i = 0; i < 64; ++i) A[i] = 0.0;
  for (int i = 0; i < 8; ++i) b[i] = 0.0;
  for (int i = 0; i < n; ++i) {
    double X, Y, u, v;
    obj(i, &X, &Y); img(i, &u, &v);
    X = (X - mo[0]) * so; Y = (Y - mo[1]) * so; u = (u - mi[0]) * si; v = (v - mi[1]) * si;
    // rows:  [X Y 1 0 0 0 -uX -uY] h = u ;  [0 0 0 X Y 1 -vX -vY] h = v      -> accumulate normal equations
    const double r1[8] = {X, Y, 1, 0, 0, 0, -u * X, -u * Y}, r2[8] = {0, 0, 0, X, Y, 1, -v * X, -v * Y};
    for (int a = 0; a < 8; ++a) {
      for (int c = 0; c < 8; ++c) A[a * 8 + c] += r1[a] * r1[c] + r2[a] * r2[c];
      b[a] += r1[a] * u + r2[a] * v;
    }
  }
  if (!solve_linear(A, b, 8)) return false;
  // H = Ti^-1 * Hn * To   with To = [so 0 -so*mo0; 0 so -so*mo1; 0 0 1], Ti likewise
  const double Hn[9] = {b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], 1.0};
  double M[9];  // Hn * To
  for (int r = 0; r < 3; ++r) {
    M[r * 3 + 0] = Hn[r * 3 + 0] * so;
    M[r * 3 + 1] = Hn[r * 3 + 1] * so;
    M[r * 3 + 2] = Hn[r * 3 + 2] - so * (Hn[r * 3 + 0] * mo[0] + Hn[r * 3 + 1] * mo[1]);
  }
  double H[9];  // Ti^-1 = [1/si 0 mi0; 0 1/si mi1; 0 0 1]
  for (int c = 0; c < 3; ++c) {
    H[0 * 3 + c] = M[0 * 3 + c] / si + mi[0] * M[2 * 3 + c];
    H[1 * 3 + c] = M[1 * 3 + c] / si + mi[1] * M[2 * 3 + c];
    H[2 * 3 + c] = M[2 * 3 + c];
  }
  // ---- 2. pose from the homography ----
  double h1[3] = {H[0], H[3], H[6]}, h2[3] = {H[1], H[4], H[7]}, h3[3] = {H[2], H[5], H[8]};
  const double n1 = sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]), n2 = sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
  if (!(n1 > 0.0) || !(n2 > 0.0)) return false;
  double lam = 2.0 / (n1 + n2);
  if (h3[2] * lam < 0.0) lam = -lam;  // the board is in front of the camera
  double r1[3], r2[3], r3[3];
  for (int k = 0; k < 3; ++k) { r1[k] = h1[k] / n1 * (lam < 0 ? -1.0 : 1.0); t[k] = h3[k] * lam; }
  // Gram-Schmidt for r2, then r3 = r1 x r2 (the refinement below removes the residual bias)
  double d = 0.0;
  for (int k = 0; k < 3; ++k) { r2[k] = h2[k] * lam; d += r1[k] * r2[k]; }
  double nn = 0.0;
  for (int k = 0; k < 3; ++k) { r2[k] -= d * r1[k]; nn += r2[k] * r2[k]; }
  if (!(nn > 0.0)) return false;
  nn = sqrt(nn);
  for (int k = 0; k < 3; ++k) r2[k] /= nn;
  cross3(r1, r2, r3);
  for (int k = 0; k < 3; ++k) { R[k * 3 + 0] = r1[k]; R[k * 3 + 1] = r2[k]; R[k * 3 + 2] = r3[k]; }
  // ---- 3. Levenberg-Marquardt refinement of sum |pi(R X + t) - u|^2 ----
  double lambda = 1e-3, cost = -1.0;
  for (int it = 0; it < 100; ++it) {
    double JtJ[36], Jtr[6], c_now = 0.0;
    for (int i = 0; i < 36; ++i) JtJ[i] = 0.0;
    for (int i = 0; i < 6; ++i) Jtr[i] = 0.0;
    for (int i = 0; i < n; ++i) {
      double X, Y, u, v;
      obj(i, &X, &Y); img(i, &u, &v);
      const double q[3] = {R[0] * X + R[1] * Y, R[3] * X + R[4] * Y, R[6] * X + R[7] * Y};  // R (X, Y, 0)
      const double pc[3] = {q[0] + t[0], q[1] + t[1], q[2] + t[2]};
      const double iz = 1.0 / pc[2];
      const double ru = pc[0] * iz - u, rv = pc[1] * iz - v;
      c_now += ru * ru + rv * rv;
      // d(u,v)/d pc
      const double a0[3] = {iz, 0.0, -pc[0] * iz * iz}, a1[3] = {0.0, iz, -pc[1] * iz * iz};
      // left perturbation R <- exp([dtheta]x) R moves pc by dtheta x q = -[q]x dtheta, so for a row a = d(u)/d(pc):
      // a^T (-[q]x) = (q x a)^T
      double Ju[6], Jv[6];
      Ju[0] = q[1] * a0[2] - q[2] * a0[1]; Ju[1] = q[2] * a0[0] - q[0] * a0[2]; Ju[2] = q[0] * a0[1] - q[1] * a0[0];
      Jv[0] = q[1] * a1[2] - q[2] * a1[1]; Jv[1] = q[2] * a1[0] - q[0] * a1[2]; Jv[2] = q[0] * a1[1] - q[1] * a1[0];
      for (int k = 0; k < 3; ++k) { Ju[3 + k] = a0[k]; Jv[3 + k] = a1[k]; }
      for (int a = 0; a < 6; ++a) {
        for (int c2 = a; c2 < 6; ++c2) JtJ[a * 6 + c2] += Ju[a] * Ju[c2] + Jv[a] * Jv[c2];
        Jtr[a] += Ju[a] * ru + Jv[a] * rv;
      }
    }
    if (cost < 0.0) cost = c_now;
    bool stepped = false;
    for (int tries = 0; tries < 12 && !stepped; ++tries) {
      double Ad[36], y[6], rhs[6];
      for (int a = 0; a < 6; ++a) {
        for (int c2 = 0; c2 < 6; ++c2) Ad[a * 6 + c2] = (a <= c2) ? JtJ[a * 6 + c2] : JtJ[c2 * 6 + a];
        Ad[a * 6 + a] *= (1.0 + lambda);
        rhs[a] = -Jtr[a];
      }
      if (!chol6_solve(Ad, rhs, y)) { lambda *= 10.0; continue; }
      // candidate: R' = exp([dtheta]x) R (Rodrigues), t' = t + dt
      const double th2 = y[0] * y[0] + y[1] * y[1] + y[2] * y[2], th = sqrt(th2);
      double E[9];
      {
        const double a = th < 1e-12 ? 1.0 : sin(th) / th, bq = th < 1e-12 ? 0.5 : (1.0 - cos(th)) / th2;
        const double K[9] = {0, -y[2], y[1], y[2], 0, -y[0], -y[1], y[0], 0};
        for (int r = 0; r < 3; ++r)
          for (int c2 = 0; c2 < 3; ++c2) {
            double kk = 0.0;
            for (int m = 0; m < 3; ++m) kk += K[r * 3 + m] * K[m * 3 + c2];
            E[r * 3 + c2] = (r == c2 ? 1.0 : 0.0) + a * K[r * 3 + c2] + bq * kk;
          }
      }
      double Rn[9], tn[3];
      for (int r = 0; r < 3; ++r) {
        for (int c2 = 0; c2 < 3; ++c2) Rn[r * 3 + c2] = E[r * 3] * R[c2] + E[r * 3 + 1] * R[3 + c2] + E[r * 3 + 2] * R[6 + c2];
        tn[r] = t[r] + y[3 + r];
      }
      double c_new = 0.0;
      for (int i = 0; i < n; ++i) {
        double X, Y, u, v;
        obj(i, &X, &Y); img(i, &u, &v);
        const double pz = Rn[6] * X + Rn[7] * Y + tn[2];
        const double ru = (Rn[0] * X + Rn[1] * Y + tn[0]) / pz - u, rv = (Rn[3] * X + Rn[4] * Y + tn[1]) / pz - v;
        c_new += ru * ru + rv * rv;
      }
      if (c_new <= c_now) {
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
        for (int k = 0; k < 3; ++k) t[k] = tn[k];
        lambda = lambda > 1e-12 ? lambda * 0.1 : lambda;
        stepped = true;
        const double dn = sqrt(th2 + y[3] * y[3] + y[4] * y[4] + y[5] * y[5]);
        if (dn < 1e-13 || c_now - c_new <= 1e-16 * c_now) return true;
      } else {
        lambda *= 10.0;
      }
    }
    if (!stepped) return true;  // no further descent possible: at the minimum to rounding
  }
  return true;
}

// All grid corners of a board at pose fp project inside the image (the tag detector needs the whole grid in view).
CLC_HD bool camera_board_in_view(const CameraDesc& cam, int width, int height, const double* fp) {
  double R[9];
  quat_to_rot(fp, R);
  const int n = grid_num_corners(cam);
  for (int i = 0; i < n; ++i) {
    double X, Y, u, v;
    grid_corner(cam, i, &X, &Y);
    const double P[3] = {R[0] * X + R[1] * Y + fp[4], R[3] * X + R[4] * Y + fp[5], R[6] * X + R[7] * Y + fp[6]};
    if (!(P[2] > 0.05)) return false;
    camera_project(cam, P, &u, &v);
    if (!(u >= 0.0 && u < (double)width && v >= 0.0 && v < (double)height)) return false;
  }
  return true;
}

// Board pose draw of the generator when a camera model is active: the accept/redraw rule of gen_frame_pose plus
// "the whole grid is in the image" (up to 512 attempts; the last draw is kept if none passes and false is returned).
CLC_HD bool gen_frame_pose_camera(const CameraDesc& cam, int width, int height, uint64_t seed, int64_t frame, bool with_edges,
                                  double* fp) {
  for (int attempt = 0; attempt < 512; ++attempt) {
    gen_draw_pose(seed, frame, attempt, fp);
    double nl[3], dl, a, b;
    gen_plane_laser(fp, nl, &dl);
    bool ok = gen_window(nl, dl, &a, &b);
    if (ok && with_edges) {
      double ep[6];
      ok = gen_edge_points(fp, ep);
    }
    if (ok) ok = camera_board_in_view(cam, width, height, fp);
    if (ok) return true;
  }
  return false;
}

// The whole chain for one frame: true pose fp_true (qx qy qz qw tx ty tz) -> estimated pose fp_est.  With
// cam.model == kCameraNone the pose is passed through unchanged.  false if the board leaves the valid camera domain.
CLC_HD bool camera_estimate_pose(const CameraDesc& cam, uint64_t seed, int64_t frame, const double* fp_true, double* fp_est,
                                 float* scratch_uv /* [2 * corners] */) {
  if (cam.model == kCameraNone) {
    for (int k = 0; k < 7; ++k) fp_est[k] = fp_true[k];
    return true;
  }
  double R[9];
  quat_to_rot(fp_true, R);
  const int n = grid_num_corners(cam);
  for (int i = 0; i < n; ++i) {
    double X, Y;
    grid_corner(cam, i, &X, &Y);
    const double P[3] = {R[0] * X + R[1] * Y + fp_true[4], R[3] * X + R[4] * Y + fp_true[5], R[6] * X + R[7] * Y + fp_true[6]};
    if (!(P[2] > 1e-6)) return false;
    double u, v, nu, nv, xn, yn;
    camera_project(cam, P, &u, &v);
    pixel_noise(seed, cam.pixel_sigma, frame, i, &nu, &nv);
    camera_lift_normalised(cam, u + nu, v + nv, &xn, &yn);
    scratch_uv[2 * i] = (float)xn;  // cv::Point2f
    scratch_uv[2 * i + 1] = (float)yn;
  }
  double Re[9], te[3];
  auto obj = [&](int i, double* X, double* Y) {
    double x, y;
    grid_corner(cam, i, &x, &y);
    *X = (double)(float)x;  // cv::Point3f
    *Y = (double)(float)y;
  };
  auto img = [&](int i, double* u, double* v) {
    *u = (double)scratch_uv[2 * i];
    *v = (double)scratch_uv[2 * i + 1];
  };
  if (!pnp_planar(n, obj, img, Re, te)) return false;
  rot_to_quat(Re, fp_est);
  fp_est[4] = te[0]; fp_est[5] = te[1]; fp_est[6] = te[2];
  return true;
}

// ---- pose of the board from its detected tag corners (reference src/calcCamPose.cpp:270-294 calcCamPose after
//      FindTargetCorner, and :211-236 EstimatePose) --------------------------------------------------------------------
// One frame: n_det detected tags, ascending id (:78-79), corners[n_det*4*2] float pixel coordinates in the detector's
// corner order (:64-70).  Every corner is undistorted onto the normalised image plane (liftProjective, x/z y/z, stored as
// cv::Point2f), paired with its kalibr-grid object point (cv::Point3f, :114-136; a single tag of id 0 is the APRILTAG
// pattern :160-178), and solvePnP with identity intrinsics gives T_cw; the function returns T_wc = T_cw^-1 (:229-230)
// as pose_wc = (qx, qy, qz, qw, x, y, z) -- what kalibratag_detector_node writes to apriltag_pose.txt.
// false (and the identity pose, :216-220) for fewer than 4 points, a tag id outside the grid, or a failed PnP.
CLC_HD bool estimate_pose_from_detections(const CameraDesc& cam, int n_det, const int* tag_ids, const float* corners,
                                          float* lifted /* scratch [n_det*8] */, double* pose_wc) {
  pose_wc[0] = pose_wc[1] = pose_wc[2] = 0.0;
  pose_wc[3] = 1.0;
  pose_wc[4] = pose_wc[5] = pose_wc[6] = 0.0;
  const int n = 4 * n_det;
  if (n < 4) return false;
  const int n_tags = cam.grid_rows * cam.grid_cols;
  for (int d = 0; d < n_det; ++d)
    if (tag_ids[d] < 0 || tag_ids[d] >= n_tags) return false;
  for (int i = 0; i < n; ++i) {
    double xn, yn;
    camera_lift_normalised(cam, (double)corners[2 * i], (double)corners[2 * i + 1], &xn, &yn);
    lifted[2 * i] = (float)xn;
    lifted[2 * i + 1] = (float)yn;
  }
  auto obj = [&](int i, double* X, double* Y) {
    double x, y;
    grid_corner(cam, tag_ids[i >> 2] * 4 + (i & 3), &x, &y);
    *X = (double)(float)x;
    *Y = (double)(float)y;
  };
  auto img = [&](int i, double* u, double* v) {
    *u = (double)lifted[2 * i];
    *v = (double)lifted[2 * i + 1];
  };
  double Rcw[9], tcw[3];
  if (!pnp_planar(n, obj, img, Rcw, tcw)) return false;
  const double Rwc[9] = {Rcw[0], Rcw[3], Rcw[6], Rcw[1], Rcw[4], Rcw[7], Rcw[2], Rcw[5], Rcw[8]};
  rot_to_quat(Rwc, pose_wc);
  for (int r = 0; r < 3; ++r) pose_wc[4 + r] = -(Rwc[3 * r] * tcw[0] + Rwc[3 * r + 1] * tcw[1] + Rwc[3 * r + 2] * tcw[2]);
  return true;
}

#ifdef __CUDACC__
// one thread per frame (the PnP refinement is sequential; the batch is the parallelism)
__global__ void clc_estimate_poses_kernel(CameraDesc cam, int64_t n_frames, const int64_t* __restrict__ det_offsets,
                                          const int* __restrict__ tag_ids, const float* __restrict__ corners,
                                          float* __restrict__ lifted, double* __restrict__ pose_wc, int* __restrict__ ok) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames) return;
  const int64_t d0 = det_offsets[i], d1 = det_offsets[i + 1];
  double pose[7];
  const bool good = estimate_pose_from_detections(cam, (int)(d1 - d0), tag_ids + d0, corners + 8 * d0, lifted + 8 * d0, pose);
  for (int k = 0; k < 7; ++k) pose_wc[i * 7 + k] = pose[k];
  ok[i] = good ? 1 : 0;
}
#endif

}  // namespace clc

"""Independent numpy twin of the C oracle.  TEST INFRASTRUCTURE ONLY (see oracle/clc_oracle.h).

Written separately from clc_oracle.c, with numpy's LAPACK (lstsq / svd / solve) in place of the hand-written
dense kernels, so that a transcription error in one restatement shows up as a disagreement between the two.
Same citations: /root/reference/src/LaseCamCalCeres.cpp, src/pose_local_parameterization.cpp and the published
Ceres (<= 2.1) trust-region semantics.  PARITY UNPINNED (no reference tests exist; Ceres/Eigen absent here).
"""
from __future__ import annotations

import numpy as np


# --- Eigen restatements --------------------------------------------------------------------------------
def quat_to_rot(q):
    """Eigen QuaternionBase::toRotationMatrix; q = (x,y,z,w), not normalised."""
    x, y, z, w = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def quat_mul(a, b):
    """Eigen quaternion product, (x,y,z,w) order."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def pose_plus(x, d):
    """pose_local_parameterization.cpp:15-32."""
    q = quat_mul(x[3:7], np.array([d[3] / 2, d[4] / 2, d[5] / 2, 1.0]))
    return np.concatenate([x[:3] + d[:3], q / np.linalg.norm(q)])


def skew(p):
    """LaseCamCalCeres.cpp:35-42."""
    return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])


# --- planes ----------------------------------------------------------------------------------------------
def frame_plane(fp):
    """LaseCamCalCeres.cpp:227-231, literally: (Tctag^-1)^T (0,0,1,0)."""
    T = np.eye(4)
    T[:3, :3] = quat_to_rot(fp[:4])
    T[:3, 3] = fp[4:7]
    return np.linalg.inv(T).T @ np.array([0.0, 0.0, 1.0, 0.0])


def pi_from_ppp(x1, x2, x3):
    """utilities.cpp:267-272."""
    return np.concatenate([np.cross(x1 - x3, x2 - x3), [-x3 @ np.cross(x1, x2)]])


def edge_planes(fp):
    """LaseCamCalCeres.cpp:262-276."""
    orig = np.array([0.0265 + 0.0165, 0.0265 + 0.0165, 0.0])
    R, t = quat_to_rot(fp[:4]), fp[4:7]
    p1c, p2c, p3c = (R @ (np.array(p) - orig) + t for p in ([0, 0, 0], [0.5, 0, 0], [0, 0.5, 0]))
    z = np.zeros(3)
    return pi_from_ppp(p1c, p2c, z), pi_from_ppp(p1c, p3c, z)


# --- residual table ---------------------------------------------------------------------------------------
def residual_table(frame_pose, offsets, points, edge_points=None):
    """Per-residual (plane[4], point[3], scale) in the reference's AddResidualBlock order (:222-295)."""
    planes, pts, scales = [], [], []
    for f in range(len(frame_pose)):
        b, e = int(offsets[f]), int(offsets[f + 1])
        if e <= b:
            continue
        pl = frame_plane(frame_pose[f])
        s = 1.0 / np.sqrt(float(e - b))
        for j in range(b, e):
            planes.append(pl); pts.append(points[j]); scales.append(s)
        if edge_points is not None:
            pi1, pi2 = edge_planes(frame_pose[f])
            planes.append(pi1); pts.append(edge_points[f, 0:3]); scales.append(s)
            planes.append(pi2); pts.append(edge_points[f, 3:6]); scales.append(s)
    return np.array(planes).reshape(-1, 4), np.array(pts).reshape(-1, 3), np.array(scales)


def evaluate(table, pose7, use_loss=True, cauchy_a=0.05):
    """PointInPlaneFactor::Evaluate (:43-66) for every row + Ceres CauchyLoss / Corrector / local Jacobian.

    Returns cost, corrected residuals [R], corrected local Jacobian [R,6]."""
    planes, pts, s = table
    R = quat_to_rot(pose7[3:7])
    t = pose7[:3]
    n = planes[:, :3]
    pc = pts @ R.T + t
    r = s * (np.einsum("ij,ij->i", n, pc) + planes[:, 3])
    J = np.empty((len(r), 6))
    J[:, :3] = s[:, None] * n
    # n^T (-R skew(p)) = (p x R^T n)^T
    m = n @ R
    J[:, 3:] = s[:, None] * np.cross(pts, m)
    if not use_loss:
        return 0.5 * float(r @ r), r, J
    b = (cauchy_a * s) ** 2
    c = 1.0 / b
    summ = 1.0 + (r * r) * c
    inv = 1.0 / summ
    rho0 = b * np.log(summ)
    rho1 = np.maximum(np.finfo(float).tiny, inv)
    sq = np.sqrt(rho1)
    return 0.5 * float(np.sum(rho0)), r * sq, J * sq[:, None]


def gradient_max_norm(x, g):
    return float(np.max(np.abs(x - pose_plus(x, -g))))


def solve(table, pose7, use_loss=True, cauchy_a=0.05, max_num_iterations=100, verbose=False):
    """Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR with the Ceres defaults
    (see SURVEY.md section 8(c) / oracle/clc_oracle.c for the line-by-line restatement)."""
    ftol, gtol, ptol = 1e-6, 1e-10, 1e-8
    radius, dec = 1e4, 2.0
    x = np.array(pose7, dtype=float)
    x_norm = np.linalg.norm(x)
    cost, r, J = evaluate(table, x, use_loss, cauchy_a)
    g = J.T @ r
    scale = 1.0 / (1.0 + np.sqrt(np.sum(J * J, axis=0)))
    J = J * scale
    trace = [dict(iteration=0, cost=cost, ok=True, gmax=gradient_max_norm(x, g), radius=radius)]
    reuse, invalid, diag = False, 0, None
    term = None
    while True:
        last = trace[-1]
        if last["iteration"] >= max_num_iterations:
            term = "NO_CONVERGENCE"; break
        if last["ok"] and last["gmax"] <= gtol:
            term = "CONVERGENCE_GRADIENT"; break
        if radius <= 1e-32:
            term = "CONVERGENCE_MIN_RADIUS"; break
        it = last["iteration"] + 1
        if not reuse:
            diag = np.clip(np.sum(J * J, axis=0), 1e-6, 1e32)
        D = np.sqrt(diag / radius)
        A = np.vstack([J, np.diag(D)])
        rhs = np.concatenate([r, np.zeros(6)])
        y = np.linalg.lstsq(A, rhs, rcond=None)[0]
        step = -y
        reuse = True
        mr = J @ step
        model_change = -float(mr @ (r + mr / 2.0))
        if not (np.all(np.isfinite(step)) and model_change > 0.0):
            invalid += 1
            if invalid >= 5:
                term = "FAILURE"; break
            radius /= dec; dec *= 2.0
            trace.append(dict(iteration=it, cost=cost, ok=False, gmax=last["gmax"], radius=radius))
            continue
        invalid = 0
        cand = pose_plus(x, step * scale)
        cand_cost, _, _ = evaluate(table, cand, use_loss, cauchy_a)
        step_norm = np.linalg.norm(x - cand)
        if step_norm <= ptol * (x_norm + ptol):
            term = "CONVERGENCE_PARAMETER"; break
        change = cost - cand_cost
        if abs(change) <= ftol * cost:
            term = "CONVERGENCE_FUNCTION"; break
        rho = change / model_change
        if rho > 1e-3:
            x = cand
            x_norm = np.linalg.norm(x)
            cost, r, J = evaluate(table, x, use_loss, cauchy_a)
            g = J.T @ r
            J = J * scale
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            dec, reuse = 2.0, False
            trace.append(dict(iteration=it, cost=cost, ok=True, gmax=gradient_max_norm(x, g), radius=radius,
                              rho=rho, step_norm=step_norm))
        else:
            radius /= dec; dec *= 2.0
            trace.append(dict(iteration=it, cost=cand_cost, ok=False, gmax=0.0, radius=radius, rho=rho,
                              step_norm=step_norm))
        if verbose:
            print(trace[-1])
    return x, term, trace


def information(frame_pose, offsets, points, pose7):
    """Analysis tail (:318-381): no loss, no edge residuals."""
    table = residual_table(frame_pose, offsets, points, None)
    _, r, J = evaluate(table, pose7, use_loss=False)
    H = J.T @ J
    return H, -J.T @ r, float(r @ r), np.linalg.svd(H, compute_uv=False)


def closed_form(frame_pose, offsets, points):
    """CamLaserCalClosedSolution (:112-203), literally with a dense A."""
    rows, rhs = [], []
    for f in range(len(frame_pose)):
        pl = frame_plane(frame_pose[f])
        for j in range(int(offsets[f]), int(offsets[f + 1])):
            bar = np.array([points[j, 0], points[j, 1], 1.0])
            rows.append(np.concatenate([pl[:3] * bar[0], pl[:3] * bar[1], pl[:3] * bar[2]]))
            rhs.append(-pl[3])
    A, b = np.array(rows), np.array(rhs)
    AtA = A.T @ A
    sv = np.linalg.svd(AtA, compute_uv=False)
    unobservable = bool(np.any(sv < 1e-10))
    H = np.linalg.solve(AtA, A.T @ b)
    h1, h2, h3 = H[0:3], H[3:6], H[6:9]
    Rcl = np.column_stack([h1, h2, np.cross(h1, h2)])
    Rlc = Rcl.T
    tlc = -Rlc @ h3
    U, _, Vt = np.linalg.svd(Rlc)
    T = np.eye(4)
    T[:3, :3] = U @ Vt
    T[:3, 3] = tlc
    return T, unobservable, AtA, A.T @ b


# --- LineFittingCeres (LaseCamCalCeres.cpp:385-433) with a generic restatement of the same Ceres loop ----------------
def trust_region_lm(evaluate_fn, plus_fn, x0, max_num_iterations, gradient_norm_fn):
    """Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR, defaults as in solve() above, for any
    residual model: evaluate_fn(x) -> (cost, corrected residuals, corrected Jacobian)."""
    ftol, gtol, ptol = 1e-6, 1e-10, 1e-8
    radius, dec = 1e4, 2.0
    x = np.array(x0, dtype=float)
    x_norm = np.linalg.norm(x)
    cost, r, J = evaluate_fn(x)
    g = J.T @ r
    scale = 1.0 / (1.0 + np.sqrt(np.sum(J * J, axis=0)))
    J = J * scale
    n = J.shape[1]
    trace = [dict(iteration=0, cost=cost, ok=True, gmax=gradient_norm_fn(x, g), radius=radius)]
    reuse, invalid, diag, term = False, 0, None, None
    while True:
        last = trace[-1]
        if last["iteration"] >= max_num_iterations:
            term = "NO_CONVERGENCE"; break
        if last["ok"] and last["gmax"] <= gtol:
            term = "CONVERGENCE_GRADIENT"; break
        if radius <= 1e-32:
            term = "CONVERGENCE_MIN_RADIUS"; break
        it = last["iteration"] + 1
        if not reuse:
            diag = np.clip(np.sum(J * J, axis=0), 1e-6, 1e32)
        A = np.vstack([J, np.diag(np.sqrt(diag / radius))])
        y = np.linalg.lstsq(A, np.concatenate([r, np.zeros(n)]), rcond=None)[0]
        step = -y
        reuse = True
        mr = J @ step
        model_change = -float(mr @ (r + mr / 2.0))
        if not (np.all(np.isfinite(step)) and model_change > 0.0):
            invalid += 1
            if invalid >= 5:
                term = "FAILURE"; break
            radius /= dec; dec *= 2.0
            trace.append(dict(iteration=it, cost=cost, ok=False, gmax=last["gmax"], radius=radius))
            continue
        invalid = 0
        cand = plus_fn(x, step * scale)
        cand_cost = evaluate_fn(cand)[0]
        if np.linalg.norm(x - cand) <= ptol * (x_norm + ptol):
            term = "CONVERGENCE_PARAMETER"; break
        change = cost - cand_cost
        if abs(change) <= ftol * cost:
            term = "CONVERGENCE_FUNCTION"; break
        rho = change / model_change
        if rho > 1e-3:
            x = cand
            x_norm = np.linalg.norm(x)
            cost, r, J = evaluate_fn(x)
            g = J.T @ r
            J = J * scale
            radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3))
            dec, reuse = 2.0, False
            trace.append(dict(iteration=it, cost=cost, ok=True, gmax=gradient_norm_fn(x, g), radius=radius))
        else:
            radius /= dec; dec *= 2.0
            trace.append(dict(iteration=it, cost=cand_cost, ok=False, gmax=0.0, radius=radius))
    return x, term, trace


def line_fit(points, line0=(0.0, 0.0), max_num_iterations=10, a=0.05):
    """residual m0 x + m1 y + 1 (:391), CauchyLoss(0.05) (:416), max_num_iterations = 10 (:425)."""
    xy = np.asarray(points, dtype=float)[:, :2]

    def ev(m):
        res = xy @ m + 1.0
        summ = 1.0 + res * res / (a * a)
        sq = np.sqrt(np.maximum(np.finfo(float).tiny, 1.0 / summ))
        return 0.5 * float(np.sum(a * a * np.log(summ))), res * sq, xy * sq[:, None]

    return trust_region_lm(ev, lambda x, d: x + d, np.array(line0, dtype=float), max_num_iterations,
                           lambda x, g: float(np.max(np.abs(g))))


# --- scan preparation (src/utilities.cpp:181-215, src/selectScanPoints.cpp:17-190), literal python loops -----------------
def scan_to_points(ranges, angle_min, angle_increment, range_min):
    r = np.asarray(ranges, dtype=np.float32)
    ang = angle_min + np.arange(len(r), dtype=float) * angle_increment
    ok = (r < 30.0) & (r >= range_min)
    pts = np.zeros((len(r), 3))
    pts[:, 0] = np.where(ok, r.astype(float) * np.cos(ang), 1000.0)
    pts[:, 1] = np.where(ok, r.astype(float) * np.sin(ang), 1000.0)
    return pts


def auto_get_line_pts(points):
    n = len(points)
    if n == 0:
        return None
    nrm = lambda i: float(np.hypot(points[i][0], points[i][1]))  # noqa: E731
    idc = n // 2
    delta = int(80 / 0.3)
    id_left, id_right = min(idc + delta, n - 1), max(idc - delta, 0)
    segs = []
    cur, nxt, new_seg, seg = id_right, id_right + 3, True, [0, 0]
    for _ in range(id_right, id_left - 3, 3):
        if new_seg:
            seg, new_seg = [cur, nxt], False
        d1, d2 = nrm(cur), nrm(nxt)
        if d1 < 100 and d2 < 100:
            if abs(d1 - d2) < 0.05:
                seg[1] = nxt
            else:
                new_seg = True
                dist = float(np.hypot(points[seg[0]][0] - points[seg[1]][0], points[seg[0]][1] - points[seg[1]][1]))
                if dist > 0.2 and nrm(seg[0]) < 2 and nrm(seg[1]) < 2 and seg[1] - seg[0] > 50:
                    segs.append(list(seg))
            cur, nxt = nxt, nxt + 3
        else:
            if d1 > 100:
                cur = nxt
            nxt += 3
    out = []
    for s0, e0 in segs:
        s, e = s0, e0
        for j in (1, 2, 3):
            if e0 + j < n and abs(nrm(e0) - nrm(e0 + j)) < 0.05:
                e = e0 + j
        for j in (-1, -2, -3):
            if s0 + j >= 0 and abs(nrm(s0) - nrm(s0 + j)) < 0.05:
                s = s0 + j
        out.append((s, e))
    if not out:
        return None
    best = max(range(len(out)), key=lambda k: (out[k][1] - out[k][0], -k))
    return out[best]


# --- camera measurement chain: board corners -> pixels -> normalised points -> PnP ---------------
# camodocal camera models restated literally (camera_models/src/PinholeCamera.cc, EquidistantCamera.cc); the equidistant
# back-projection takes the smallest non-negative real root from the companion-matrix eigenvalues (np.roots), exactly
# as EquidistantCamera.cc:632-734 does.  The PnP is OpenCV itself (cv2.solvePnP), the dependency the reference calls.
def pinhole_distortion(k, x, y):
    k1, k2, p1, p2 = k[4:8]
    rho2 = x * x + y * y
    rad = k1 * rho2 + k2 * rho2 * rho2
    return x * rad + 2.0 * p1 * x * y + p2 * (rho2 + 2.0 * x * x), y * rad + 2.0 * p2 * x * y + p1 * (rho2 + 2.0 * y * y)


def camera_project(model, k, P):
    """Camera::spaceToPlane: model 1 = pinhole radtan (fx fy cx cy k1 k2 p1 p2), 2 = equidistant (mu mv u0 v0 k2..k5)."""
    P = np.asarray(P, dtype=float)
    if model == 2:
        theta = np.arccos(P[2] / np.linalg.norm(P))
        phi = np.arctan2(P[1], P[0])
        r = theta + k[4] * theta**3 + k[5] * theta**5 + k[6] * theta**7 + k[7] * theta**9
        return np.array([k[0] * r * np.cos(phi) + k[2], k[1] * r * np.sin(phi) + k[3]])
    x, y = P[0] / P[2], P[1] / P[2]
    if any(k[4:8]):
        dx, dy = pinhole_distortion(k, x, y)
        x, y = x + dx, y + dy
    return np.array([k[0] * x + k[2], k[1] * y + k[3]])


def camera_lift_normalised(model, k, uv):
    """Camera::liftProjective followed by the division by z of src/calcCamPose.cpp:290-291."""
    mx, my = (uv[0] - k[2]) / k[0], (uv[1] - k[3]) / k[1]
    if model == 2:
        rn = float(np.hypot(mx, my))
        phi = 0.0 if rn < 1e-10 else float(np.arctan2(my, mx))
        coeffs = {1: 1.0, 3: k[4], 5: k[5], 7: k[6], 9: k[7]}
        npow = 9
        for kk in (k[7], k[6], k[5], k[4]):
            if kk == 0.0:
                npow -= 2
        if npow == 1:
            theta = rn
        else:
            poly = np.zeros(npow + 1)  # highest power first for np.roots
            for pw, c in coeffs.items():
                if pw <= npow:
                    poly[npow - pw] = c
            poly[npow] = -rn
            cands = []
            for rt in np.roots(poly):
                if abs(rt.imag) > 1e-10 or rt.real < -1e-10:
                    continue
                cands.append(max(rt.real, 0.0))
            theta = min(cands) if cands else rn
        return np.array([np.tan(theta) * np.cos(phi), np.tan(theta) * np.sin(phi)])
    xu, yu = mx, my
    if any(k[4:8]):
        dx, dy = pinhole_distortion(k, mx, my)
        xu, yu = mx - dx, my - dy
        for _ in range(7):
            dx, dy = pinhole_distortion(k, xu, yu)
            xu, yu = mx - dx, my - dy
    return np.array([xu, yu])


def grid_corners(rows, cols, tag_size, tag_spacing):
    """kalibr april grid, reference src/calcCamPose.cpp:107-136: [4*rows*cols, 2]."""
    pitch = tag_size * (1.0 + tag_spacing)
    out = []
    for tag in range(rows * cols):
        r, c = divmod(tag, cols)
        x0, y0 = pitch * c, pitch * r
        out += [(x0, y0), (x0 + tag_size, y0), (x0 + tag_size, y0 + tag_size), (x0, y0 + tag_size)]
    return np.array(out)


def estimate_pose_cv(model, k, corners_xy, fp_true, pixel_noise=None):
    """The reference's chain with OpenCV's solvePnP (src/calcCamPose.cpp:211-236,279-292): returns (R_ca, t_ca) estimated."""
    import cv2

    R = quat_to_rot(fp_true[:4])
    t = np.asarray(fp_true[4:7], dtype=float)
    un = []
    for i, (X, Y) in enumerate(corners_xy):
        uv = camera_project(model, k, R @ np.array([X, Y, 0.0]) + t)
        if pixel_noise is not None:
            uv = uv + pixel_noise[i]
        un.append(camera_lift_normalised(model, k, uv))
    p3 = np.c_[corners_xy, np.zeros(len(corners_xy))].astype(np.float32)  # cv::Point3f
    p2 = np.array(un, dtype=np.float32)                                    # cv::Point2f
    ok, rvec, tvec = cv2.solvePnP(p3, p2, np.eye(3, dtype=np.float32), np.zeros((1, 5), dtype=np.float32))
    Rm, _ = cv2.Rodrigues(rvec)
    return Rm, tvec.ravel(), p2

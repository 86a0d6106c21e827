"""Regenerates the golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

No golden vectors exist upstream, and the reference cannot run here (no Ceres / Eigen /
ROS), so these fixtures are produced by the two independent restatements of the reference algorithm:
  * factor_kat.json   : residual + 1x6 Jacobian of PointInPlaneFactor for hand-checkable inputs, derived with sympy
                        by differentiating n.(R(q (x) dq(dtheta)) p + t + dt) + d symbolically (independent of both
                        oracles' analytic Jacobians).
  * preprocessing.json: eight LaserScan-like range arrays (self-contained float32 inputs) -> board segment indices
                        (AutoGetLinePts) and the robust line fit of the segment (LineFittingCeres), numpy twin.
  * camera_chain.json : both camera models: 3-D points -> pixels, pixels -> normalised image plane (np.roots for the
                        Kannala-Brandt back-projection), and tag detections -> T_wc by OpenCV's solvePnP itself
                        (the library the reference calls at src/calcCamPose.cpp:225).
  * config1_seed*.json: BASELINE config 1 (50 frames x 180 beams, faithful ragged generator, 1 cm range noise):
                        generator outputs, (cost, H, g) at the identity start and at ground truth, the full LM
                        trajectory and result, closed form, analysis tail -- computed with the numpy twin
                        (oracle/oracle_np.py, LAPACK QR) so that the C oracle is checked against them.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def factor_kat():
    import sympy as sp

    d = sp.symbols("d0:6", real=True)

    def qmul(a, b):  # (x,y,z,w)
        ax, ay, az, aw = a
        bx, by, bz, bw = b
        return (aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz)

    def rot(q):
        x, y, z, w = q
        return sp.Matrix([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])

    s2 = sp.sqrt(2) / 2
    half = sp.Rational(1, 2)
    cases = [
        # plane (n,d), point, scale, pose (t, q)
        dict(plane=(0, 0, 1, -2), pt=(1, 2, 0), scale=1, t=(0, 0, 0), q=(0, 0, 0, 1)),
        dict(plane=(1, 0, 0, sp.Rational(-1, 2)), pt=(3, -1, 0), scale=half, t=(sp.Rational(1, 10), 0, 0), q=(0, 0, s2, s2)),
        dict(plane=(0, 1, 0, 1), pt=(2, 1, sp.Rational(1, 2)), scale=sp.Rational(1, 3), t=(0, sp.Rational(1, 5), 0), q=(s2, 0, 0, s2)),
        dict(plane=(sp.Rational(3, 5), 0, sp.Rational(4, 5), -1), pt=(1, 1, 1), scale=sp.Rational(1, 4), t=(1, 2, 3), q=(0, s2, 0, s2)),
        dict(plane=(sp.Rational(2, 3), sp.Rational(-1, 3), sp.Rational(2, 3), sp.Rational(3, 10)), pt=(-2, sp.Rational(1, 2), sp.Rational(1, 4)),
             scale=sp.Rational(1, 7), t=(sp.Rational(1, 5), sp.Rational(3, 10), sp.Rational(-1, 10)), q=(half, -half, half, half)),
    ]
    out = []
    for c in cases:
        q = c["q"]
        dq = (d[3] / 2, d[4] / 2, d[5] / 2, 1)
        qq = qmul(q, dq)
        nrm = sp.sqrt(sum(v * v for v in qq))
        qq = tuple(v / nrm for v in qq)
        p = sp.Matrix(c["pt"])
        t = sp.Matrix(c["t"]) + sp.Matrix(d[:3])
        n = sp.Matrix(c["plane"][:3])
        r = c["scale"] * ((n.T * (rot(qq) * p + t))[0] + c["plane"][3])
        zero = {v: 0 for v in d}
        r0 = sp.N(r.subs(zero), 30)
        J = [sp.N(sp.diff(r, v).subs(zero), 30) for v in d]
        out.append(dict(plane=[float(v) for v in c["plane"]], pt=[float(v) for v in c["pt"]], scale=float(c["scale"]),
                        pose7=[float(v) for v in c["t"]] + [float(sp.N(v, 30)) for v in q], r=float(r0), J=[float(v) for v in J]))
    return out


def config1(seed):
    from oracle import oracle as O
    from oracle import oracle_np as N

    p = O.generate(50, 180, seed=seed, sigma=0.01)
    _, gt = O.ground_truth()
    tab = N.residual_table(p.frame_pose, p.offsets, p.points, None)
    x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
    out = dict(seed=seed, n_frames=50, beams=180, sigma=0.01, n_points=int(p.n_points), offsets=p.offsets.tolist(),
               frame_pose_first3=p.frame_pose[:3].tolist(), point_first=p.points[0].tolist(), point_last=p.points[-1].tolist(),
               points_checksum=float(np.sum(p.points * np.arange(1, 3 * p.n_points + 1).reshape(-1, 3) % 7)))
    for name, x in (("identity", x0), ("ground_truth", gt)):
        cost, r, J = N.evaluate(tab, x)
        out["eval_" + name] = dict(pose7=x.tolist(), cost=cost, H=(J.T @ J).tolist(), g=(J.T @ r).tolist())
    x, term, trace = N.solve(tab, x0)
    out["solve"] = dict(termination=term, pose7=x.tolist(), costs=[t["cost"] for t in trace],
                        accepted=[bool(t["ok"]) for t in trace], radius=[t["radius"] for t in trace])
    T, un, AtA, Atb = N.closed_form(p.frame_pose, p.offsets, p.points)
    out["closed_form"] = dict(Tlc=T.tolist(), unobservable=un, AtA_trace=float(np.trace(AtA)), Atb=Atb.tolist())
    H, b, chi, sv = N.information(p.frame_pose, p.offsets, p.points, x)
    out["information"] = dict(H=H.tolist(), b=b.tolist(), chi=chi, singular_values=sv.tolist())
    return out


def preprocessing():
    from oracle import oracle_np as N

    rng = np.random.default_rng(17)
    n_beams = 1081
    a0, inc = -2.356, 4.712 / (n_beams - 1)
    ang = a0 + np.arange(n_beams) * inc
    scans = []
    for k in range(8):
        r = (5 + np.sin(ang * 3 + rng.uniform(0, 6)) * 1.5 + rng.normal(size=n_beams) * 0.01).astype(np.float32)
        if k % 4 != 3:
            c, w, d = rng.uniform(-0.6, 0.6), rng.uniform(0.15, 0.35), rng.uniform(0.6, 1.5)
            m = np.abs(ang - c) < w
            r[m] = (d / np.cos(ang[m] - c) + rng.normal(size=int(m.sum())) * 0.003).astype(np.float32)
        r[rng.random(n_beams) < 0.01] = np.inf
        pts = N.scan_to_points(r, a0, inc, 0.05)
        seg = N.auto_get_line_pts(pts)
        entry = dict(ranges=[float(v) if np.isfinite(v) else "inf" for v in r], segment=None if seg is None else list(seg))
        if seg is not None:
            line, term, trace = N.line_fit(pts[seg[0]:seg[1] + 1])
            entry["line"] = [float(v) for v in line]
            entry["line_termination"] = term
            entry["line_iterations"] = len(trace)
            entry["point_first"] = pts[seg[0]].tolist()
        scans.append(entry)
    return dict(angle_min=a0, angle_increment=inc, range_min=0.05, scans=scans)


def camera_chain():
    import cv2

    from oracle import oracle_np as N

    rng = np.random.default_rng(23)
    out = {}
    models = {"radtan": (1, [367.05, 366.94, 368.72, 241.14, -0.28, 0.07, 0.0003, -0.0002]),
              "equi": (2, [363.0, 363.2, 370.1, 240.3, -0.013, 0.021, -0.034, 0.012])}
    grid = (6, 6, 0.055, 0.3)
    corners = N.grid_corners(*grid)[:, :2].reshape(36, 4, 2)
    for name, (model, k) in models.items():
        k = np.array(k)
        P = np.c_[rng.uniform(-0.8, 0.8, 12), rng.uniform(-0.6, 0.6, 12), rng.uniform(0.5, 3.0, 12)]
        P = P[np.hypot(P[:, 0], P[:, 1]) / P[:, 2] < 0.4]  # where the 8-step radtan recursion has converged
        uv = np.array([N.camera_project(model, k, p) for p in P])
        lift = np.array([N.camera_lift_normalised(model, k, q) for q in uv])
        frames = []
        for _ in range(6):
            ax = rng.normal(size=3)
            ax /= np.linalg.norm(ax)
            a = rng.uniform(0, 0.6)
            q = np.concatenate([ax * np.sin(a / 2), [np.cos(a / 2)]])
            t = np.array([rng.uniform(-0.4, 0.1), rng.uniform(-0.35, 0.05), rng.uniform(0.5, 2.0)])
            R = N.quat_to_rot(q)
            ids = np.sort(rng.choice(36, size=int(rng.integers(6, 37)), replace=False))
            pix = np.array([[N.camera_project(model, k, R @ np.array([X, Y, 0.0]) + t) for X, Y in corners[i]] for i in ids])
            pix = (pix + rng.normal(scale=0.15, size=pix.shape)).astype(np.float32)
            p2 = np.array([N.camera_lift_normalised(model, k, c.astype(float)) for c in pix.reshape(-1, 2)], dtype=np.float32)
            p3 = np.c_[corners[ids].reshape(-1, 2), np.zeros(4 * len(ids))].astype(np.float32)
            _, rvec, tvec = cv2.solvePnP(p3, p2, np.eye(3, dtype=np.float32), np.zeros((1, 5), dtype=np.float32))
            Rcw, _ = cv2.Rodrigues(rvec)
            frames.append(dict(tag_ids=ids.tolist(), corners_uv=pix.astype(float).tolist(), Rwc=Rcw.T.tolist(),
                               twc=(-Rcw.T @ tvec.ravel()).tolist()))
        out[name] = dict(model=model, intrinsics=k.tolist(), grid=list(grid), points=P.tolist(), pixels=uv.tolist(),
                         lifted=lift.tolist(), frames=frames)
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "preprocessing.json"), "w") as f:
        json.dump(preprocessing(), f)
    with open(os.path.join(HERE, "camera_chain.json"), "w") as f:
        json.dump(camera_chain(), f)
    with open(os.path.join(HERE, "factor_kat.json"), "w") as f:
        json.dump(factor_kat(), f, indent=1)
    for seed in (1, 2, 3):
        with open(os.path.join(HERE, f"config1_seed{seed}.json"), "w") as f:
            json.dump(config1(seed), f, indent=1)
    print("golden fixtures written to", HERE)

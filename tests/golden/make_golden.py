"""Regenerates the golden fixtures under tests/golden/ (run from the repo root: python tests/golden/make_golden.py).

No golden vectors exist upstream, and the reference cannot run here (no Ceres / Eigen /
ROS), so these fixtures are produced by the two independent restatements of the reference algorithm:
  * factor_kat.json   : residual + 1x6 Jacobian of PointInPlaneFactor for hand-checkable inputs, derived with sympy
                        by differentiating n.(R(q (x) dq(dtheta)) p + t + dt) + d symbolically (independent of both
                        oracles' analytic Jacobians).
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


if __name__ == "__main__":
    with open(os.path.join(HERE, "factor_kat.json"), "w") as f:
        json.dump(factor_kat(), f, indent=1)
    for seed in (1, 2, 3):
        with open(os.path.join(HERE, f"config1_seed{seed}.json"), "w") as f:
            json.dump(config1(seed), f, indent=1)
    print("golden fixtures written to", HERE)

"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python profiles/sanitizer_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200 import Problem  # noqa: E402
from camlasercalibratool_b200 import formats as fmt  # noqa: E402
from oracle import oracle as O  # noqa: E402

x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
os.environ.setdefault("CLC_PLANAR_MIN_POINTS", "0")  # the two-stream kernels also on these small problems
# round 2: the loop drivers and the one-cluster kernel (cluster barriers, distributed shared memory) on a small problem
from camlasercalibratool_b200 import LineFittingCeres  # noqa: E402

ps = O.generate(50, 180, seed=1, sigma=0.01, with_edges=False)
for small, loop in (("1", "1"), ("0", "1"), ("0", "0"), ("0", "2")):
    os.environ["CLC_SMALL_KERNEL"], os.environ["CLC_LOOP_IN_KERNEL"] = small, loop
    with Problem.from_arrays(ps.frame_pose, ps.offsets, ps.points) as g:
        c, H, gr = g.eval(x0)
        x, s, tr = g.solve(x0)
        g.information(x)
        xo, so, _ = O.solve(ps, x0)
        assert O.pose_error(x, xo)[0] < 1e-8 and s.termination == so.termination, (small, loop)
line = np.zeros(2)
LineFittingCeres(ps.points[:150], line)  # the per-scan light path
os.environ["CLC_SMALL_KERNEL"], os.environ["CLC_LOOP_IN_KERNEL"] = "0", "2"
with Problem.synthetic(300, 700, seed=2, sigma=0.01) as g:  # persistent multi-block grid: tagged-word gather + pose hand-out
    g.solve(x0)
os.environ["CLC_SMALL_KERNEL"], os.environ["CLC_LOOP_IN_KERNEL"] = "0", "1"  # below: the streaming kernels on everything
for planar in ("1", "0"):
    os.environ["CLC_PLANAR"] = planar
    for edges in (False, True):
        p = O.generate(60, 150, seed=3, sigma=0.01, exact_m=edges, with_edges=edges)
        with Problem.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points) as g:
            c, H, gr = g.eval(x0)
            rc, rH, rg = O.evaluate_normal(p, x0)
            assert abs(c - rc) <= 1e-11 * rc
            x, s, tr = g.solve(x0)
            xo, so, _ = O.solve(p, x0)
            assert O.pose_error(x, xo)[0] < 1e-8 and s.termination == so.termination
            g.information(x)
            g.closed_form()
            g.line_fit()
            g.download()
for planar in ("1", "0"):
    os.environ["CLC_PLANAR"] = planar
    with Problem.synthetic(300, 700, seed=2, sigma=0.01, with_edges=True) as g:  # ragged ends of warp ranges, several blocks
        assert g.planar == (planar == "1")
        x, s, tr = g.solve(x0)
        g.bench_eval(x, 2, flush_l2=False)
        g.set_planar_mode(1 - int(planar))  # re-partition (and z re-materialised / dropped) on a live problem
        g.eval(x)
with Problem.synthetic(40, 64, seed=5, sigma=0.01, camera="equi", pixel_sigma=0.2) as g:  # generator with the camera chain
    g.solve(x0)
    g.download_true_poses()
rng = np.random.default_rng(0)
ranges = (5 + rng.normal(size=(64, 1081)) * 0.01).astype(np.float32)
ranges[:, 500:600] = 1.0
fmt.auto_get_line_segments(ranges, -2.356, 4.712 / 1080, 0.05)
dets = [(np.arange(36, dtype=np.int32), (rng.uniform(50, 400, size=(36, 4, 2))).astype(np.float32)) for _ in range(70)]
dets.append((np.zeros(0, dtype=np.int32), np.zeros((0, 4, 2), dtype=np.float32)))
fmt.estimate_board_poses("radtan", dets)
print("sanitizer smoke ok")

"""Throughput of the widened rows through the public API, next to their CPU counterparts on the same box.
python profiles/widened_rows_timing.py  -> one JSON line.  Wall clock around synchronous library calls (host buffers in,
host results out), median of 5 after one warm-up."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from camlasercalibratool_b200 import Problem  # noqa: E402
from camlasercalibratool_b200 import formats as fmt  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import oracle_np as N  # noqa: E402


def med(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        r = fn()
        ts.append(1e3 * (time.perf_counter() - t))
    return float(np.median(ts)), r


out = {}

# ---- rank 1: LineFittingCeres, batched: 10^4 scans x 10^3 points resident on the device --------------------------------
n_scans, n_pts = 10_000, 1_000
with Problem.synthetic(n_scans, n_pts, seed=1, sigma=0.01) as g:
    ms, (lines, info) = med(lambda: g.line_fit())
    sweeps = float(info[:, 2].sum())
    d = g.download()
t0 = time.perf_counter()
n_cpu = 200
for f in range(n_cpu):
    O.line_fit(d["points"][f * n_pts:(f + 1) * n_pts])
cpu_ms = 1e3 * (time.perf_counter() - t0) / n_cpu
out["line_fit"] = {"scans": n_scans, "points_per_scan": n_pts, "gpu_ms_batch": ms, "scans_per_s": n_scans / (ms * 1e-3),
                   "lm_sweeps_total": sweeps, "streamed_GBps": sweeps * n_pts * 16 / (ms * 1e-3) / 1e9,
                   "cpu_oracle_ms_per_scan_1_thread": cpu_ms, "cpu_scans_per_s": 1e3 / cpu_ms}

# ---- rank 1, per call: LineFittingCeres(Points, Line) on ONE scan, as reference main/calibr_offline.cpp:124 calls it ------------
from camlasercalibratool_b200 import LineFittingCeres  # noqa: E402

per_call = {}
for npts in (200, 1000):
    scan = np.ascontiguousarray(d["points"][:npts])
    for _ in range(20):
        LineFittingCeres(scan, np.zeros(2))
    ts = []
    for _ in range(200):
        line = np.zeros(2)
        t0 = time.perf_counter()
        LineFittingCeres(scan, line)
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(100):
        O.line_fit(scan)
    per_call[str(npts)] = {"gpu_ms_per_call_median": 1e3 * float(np.median(ts)), "gpu_ms_per_call_min": 1e3 * float(np.min(ts)),
                           "cpu_oracle_ms_per_call": 1e3 * (time.perf_counter() - t0) / 100}
out["line_fit_per_call"] = per_call

# ---- rank 4: TranScanToPoints + AutoGetLinePts, batched: 10^4 scans x 1081 beams from host memory ----------------------
rng = np.random.default_rng(0)
n_beams = 1081
a0, inc = -2.356, 4.712 / (n_beams - 1)
ang = a0 + np.arange(n_beams) * inc
base = np.empty((64, n_beams), dtype=np.float32)
for k in range(64):
    r = (5 + np.sin(ang * 3 + rng.uniform(0, 6)) * 1.5 + rng.normal(size=n_beams) * 0.01).astype(np.float32)
    c, w, dd = rng.uniform(-0.6, 0.6), rng.uniform(0.15, 0.35), rng.uniform(0.6, 1.5)
    m = np.abs(ang - c) < w
    r[m] = (dd / np.cos(ang[m] - c) + rng.normal(size=int(m.sum())) * 0.003).astype(np.float32)
    base[k] = r
ranges = np.ascontiguousarray(np.tile(base, (n_scans // 64 + 1, 1))[:n_scans])
ms, (s, e) = med(lambda: fmt.auto_get_line_segments(ranges, a0, inc, 0.05))
t0 = time.perf_counter()
for k in range(n_cpu):
    O.auto_get_line_pts(O.scan_to_points(ranges[k], a0, inc, 0.05))
cpu_ms = 1e3 * (time.perf_counter() - t0) / n_cpu
out["scan_segments"] = {"scans": n_scans, "beams": n_beams, "gpu_ms_batch_incl_h2d": ms, "scans_per_s": n_scans / (ms * 1e-3),
                        "found": int((s >= 0).sum()), "cpu_oracle_ms_per_scan_1_thread": cpu_ms, "cpu_scans_per_s": 1e3 / cpu_ms}

# ---- rank 2 / pose side: board poses from tag detections, batched: 10^4 frames x 36 tags -------------------------------
k_equi = np.array([363.0, 363.2, 370.1, 240.3, -0.013, 0.021, -0.034, 0.012])
corners = N.grid_corners(6, 6, 0.055, 0.3)[:, :2].reshape(36, 4, 2)
dets64 = []
for _ in range(64):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = rng.uniform(0, 0.6)
    q = np.concatenate([ax * np.sin(a / 2), [np.cos(a / 2)]])
    t = np.array([rng.uniform(-0.4, 0.1), rng.uniform(-0.35, 0.05), rng.uniform(0.5, 2.0)])
    R = N.quat_to_rot(q)
    pix = np.array([[N.camera_project(2, k_equi, R @ np.array([X, Y, 0.0]) + t) for X, Y in corners[i]] for i in range(36)])
    dets64.append((np.arange(36, dtype=np.int32), (pix + rng.normal(scale=0.15, size=pix.shape)).astype(np.float32)))
n_frames = 10_000
dets = [dets64[i % 64] for i in range(n_frames)]
ms, (pose, ok) = med(lambda: fmt.estimate_board_poses("equi", dets, intrinsics=k_equi), n=3)
# the C call alone, on pre-marshalled arrays (H2D + kernel + D2H)
import ctypes as C  # noqa: E402

from camlasercalibratool_b200 import _lib  # noqa: E402

cd = _lib.CameraDesc()
cd.camera_model = 2
cd.intrinsics = (C.c_double * 8)(*[float(v) for v in k_equi])
cd.grid_rows, cd.grid_cols, cd.tag_size, cd.tag_spacing = 6, 6, 0.055, 0.3
off = (np.arange(n_frames + 1) * 36).astype(np.int64)
ids_all = np.ascontiguousarray(np.tile(np.arange(36, dtype=np.int32), n_frames))
uv_all = np.ascontiguousarray(np.concatenate([dets64[i % 64][1].reshape(-1, 8) for i in range(n_frames)]), dtype=np.float32)
pose2 = np.zeros((n_frames, 7))
ok2 = np.zeros(n_frames, dtype=np.int32)
L = _lib.load()


def c_call():
    _lib.check(L.clc_estimate_board_poses(C.byref(cd), n_frames, off.ctypes.data_as(C.POINTER(C.c_int64)),
                                          ids_all.ctypes.data_as(C.POINTER(C.c_int32)), uv_all.ctypes.data_as(C.POINTER(C.c_float)),
                                          pose2.ctypes.data_as(C.POINTER(C.c_double)), ok2.ctypes.data_as(C.POINTER(C.c_int32)), -1),
               "clc_estimate_board_poses")


ms_c, _ = med(c_call)
assert np.array_equal(pose2, pose)
cpu = None
try:
    import cv2

    t0 = time.perf_counter()
    for i in range(n_cpu):
        ids, uv = dets[i]
        p2 = np.array([N.camera_lift_normalised(2, k_equi, c.astype(float)) for c in uv.reshape(-1, 2)], dtype=np.float32)
        p3 = np.c_[corners[ids].reshape(-1, 2), np.zeros(4 * len(ids))].astype(np.float32)
        t1 = time.perf_counter()
        cv2.solvePnP(p3, p2, np.eye(3, dtype=np.float32), np.zeros((1, 5), dtype=np.float32))
        cpu = (cpu or 0.0) + (time.perf_counter() - t1)
    cpu = 1e3 * cpu / n_cpu
except ImportError:
    pass
out["board_poses"] = {"frames": n_frames, "corners_per_frame": 144, "gpu_ms_batch_incl_python_marshalling": ms,
                      "gpu_ms_batch_c_abi": ms_c, "frames_per_s": n_frames / (ms_c * 1e-3), "ok": int(ok.sum()),
                      "opencv_solvePnP_ms_per_frame_1_thread": cpu, "opencv_frames_per_s": None if cpu is None else 1e3 / cpu}
print(json.dumps(out))

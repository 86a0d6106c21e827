"""The fixtures of the widened rows: tests/golden/preprocessing.json (scan -> board segment -> robust line)
and tests/golden/camera_chain.json (camera models, pose from tag detections by OpenCV's solvePnP).  CPU: the C oracle and
the host build of the product's code against them; GPU: the library through the C ABI against them -- self-contained
inputs, no oracle involved on the GPU side."""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def scans_of(fix):
    return [np.array([np.inf if v == "inf" else v for v in s["ranges"]], dtype=np.float32) for s in fix["scans"]]


def test_oracle_against_preprocessing_fixture(oracle):
    fix = load("preprocessing.json")
    a0, inc, rmin = fix["angle_min"], fix["angle_increment"], fix["range_min"]
    found = 0
    for s, r in zip(fix["scans"], scans_of(fix)):
        pts = oracle.scan_to_points(r, a0, inc, rmin)
        seg = oracle.auto_get_line_pts(pts)
        assert (seg is None and s["segment"] is None) or list(seg) == s["segment"]
        if seg is not None:
            found += 1
            np.testing.assert_allclose(pts[seg[0]], s["point_first"], atol=1e-12)
            line, summ, tr = oracle.line_fit(pts[seg[0]:seg[1] + 1])
            np.testing.assert_allclose(line, s["line"], rtol=0, atol=1e-9)
            assert oracle.TERMINATION[summ.termination] == s["line_termination"]
    assert found == 6  # scans 3 and 7 carry no board


def test_host_build_against_camera_fixture(harness):
    fix = load("camera_chain.json")
    L = harness.L
    dp = C.POINTER(C.c_double)
    L.harness_camera_project.argtypes = [C.c_int, dp, dp, dp]
    L.harness_camera_lift.argtypes = [C.c_int, dp, dp, dp]
    L.harness_estimate_pose_from_detections.argtypes = [C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                                        C.POINTER(C.c_int32), C.POINTER(C.c_float), dp]
    from oracle import oracle_np as N

    for name, m in fix.items():
        k = np.array(m["intrinsics"])
        kp = k.ctypes.data_as(dp)
        for P, uv, xy in zip(np.array(m["points"]), np.array(m["pixels"]), np.array(m["lifted"])):
            got = np.empty(2)
            L.harness_camera_project(m["model"], kp, P.ctypes.data_as(dp), got.ctypes.data_as(dp))
            np.testing.assert_allclose(got, uv, atol=1e-9)
            L.harness_camera_lift(m["model"], kp, uv.ctypes.data_as(dp), got.ctypes.data_as(dp))
            np.testing.assert_allclose(got, xy, atol=2e-9)
        rows, cols, tag, sp = m["grid"]
        for fr in m["frames"]:
            ids = np.array(fr["tag_ids"], dtype=np.int32)
            uv = np.array(fr["corners_uv"], dtype=np.float32)
            pose = np.empty(7)
            assert L.harness_estimate_pose_from_detections(m["model"], kp, int(rows), int(cols), tag, sp, len(ids),
                                                           ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                           uv.ctypes.data_as(C.POINTER(C.c_float)), pose.ctypes.data_as(dp)) == 1
            Re = N.quat_to_rot(pose[:4])
            ang = np.arccos(np.clip((np.trace(Re.T @ np.array(fr["Rwc"])) - 1) / 2, -1, 1))
            assert ang < 3e-5 and np.linalg.norm(pose[4:] - fr["twc"]) < 1.5e-4, (name, ang)


@pytest.mark.gpu
def test_library_against_preprocessing_fixture():
    from camlasercalibratool_b200 import LineFittingCeres
    from camlasercalibratool_b200 import formats as fmt

    fix = load("preprocessing.json")
    a0, inc, rmin = fix["angle_min"], fix["angle_increment"], fix["range_min"]
    ranges = np.stack(scans_of(fix))
    s, e = fmt.auto_get_line_segments(ranges, a0, inc, rmin)
    segs = fmt.segments_from_scans(np.arange(len(ranges)) * 0.1, ranges, a0, inc, rmin)
    it = iter(segs)
    for k, sc in enumerate(fix["scans"]):
        if sc["segment"] is None:
            assert s[k] == -1 and e[k] == -1
            continue
        assert [int(s[k]), int(e[k])] == sc["segment"]
        ts, pts = next(it)
        assert abs(ts - 0.1 * k) < 1e-12
        np.testing.assert_allclose(pts[0], sc["point_first"], atol=1e-12)
        line = np.zeros(2)
        LineFittingCeres(pts, line)  # the reference's signature: Line is in/out
        np.testing.assert_allclose(line, sc["line"], rtol=0, atol=1e-9)


@pytest.mark.gpu
def test_library_against_camera_fixture():
    from camlasercalibratool_b200 import formats as fmt

    fix = load("camera_chain.json")
    for name, m in fix.items():
        dets = [(np.array(fr["tag_ids"], dtype=np.int32), np.array(fr["corners_uv"], dtype=np.float32)) for fr in m["frames"]]
        pose, ok = fmt.estimate_board_poses(name, dets, intrinsics=m["intrinsics"], grid=m["grid"])
        assert ok.all()
        for p7, fr in zip(pose, m["frames"]):
            Re = fmt.quat_to_rot(p7[:4])
            ang = np.arccos(np.clip((np.trace(Re.T @ np.array(fr["Rwc"])) - 1) / 2, -1, 1))
            assert ang < 3e-5 and np.linalg.norm(p7[4:] - fr["twc"]) < 1.5e-4, (name, ang)

"""Text formats and the ROS-free offline driver glue: reference src/utilities.cpp:6-54,
main/kalibratag_detector_node.cpp:202-236, main/calibr_offline.cpp:52-197."""
import math

import numpy as np
import pytest


def _poses_from_problem(p, fmt):
    poses = []
    for f in range(p.n_frames):
        qca, tca = p.frame_pose[f, :4], p.frame_pose[f, 4:]
        qwc = fmt.quat_inverse(qca)
        twc = -fmt.quat_to_rot(qwc) @ tca
        poses.append(fmt.CamPose(100.0 + 0.5 * f, qwc, twc))
    return poses


def test_pose_txt_round_trip_and_layout(tmp_path, oracle):
    from camlasercalibratool_b200 import formats as fmt

    p = oracle.generate(12, 20, seed=3, exact_m=True)
    poses = _poses_from_problem(p, fmt)
    path = tmp_path / "apriltag_pose.txt"
    fmt.save_cam_pose_txt(path, poses)
    lines = path.read_text().splitlines()
    assert len(lines) == 12 and len(lines[0].split()) == 11  # ts x y z qx qy qz qw roll pitch yaw
    assert lines[0].split()[0] == "100.000000000" and len(lines[0].split()[1].split(".")[1]) == 10
    path.write_text(path.read_text() + "\n\n")  # trailing blank lines are skipped (utilities.cpp:25)
    back = fmt.load_cam_pose_txt(path)
    assert len(back) == 12
    for a, b in zip(poses, back):
        assert a.timestamp == b.timestamp
        np.testing.assert_allclose(a.twc, b.twc, atol=1e-10)
        np.testing.assert_allclose(a.qwc, b.qwc, atol=1e-10)


def test_euler_angles_match_scipy():
    from scipy.spatial.transform import Rotation

    from camlasercalibratool_b200 import formats as fmt

    rng = np.random.default_rng(0)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        r, p, y = fmt.to_euler_angles(q)
        yz, py, rx = Rotation.from_quat(q).as_euler("ZYX")  # yaw about Z, pitch about Y, roll about X
        np.testing.assert_allclose([r, p, y], [rx, py, yz], atol=1e-12)
    assert fmt.to_euler_angles(np.array([0, math.sin(math.pi / 4), 0, math.cos(math.pi / 4)]))[1] == pytest.approx(math.pi / 2)


def test_result_yaml_is_readable_by_opencv(tmp_path, oracle):
    from camlasercalibratool_b200 import formats as fmt

    gtT, _ = oracle.ground_truth()
    path = str(tmp_path / "result.yaml")
    rpy = fmt.write_result_yaml(path, gtT)
    mine = fmt.read_result_yaml(path)
    np.testing.assert_array_equal(mine["extrinsicTlc"], gtT)
    np.testing.assert_allclose(mine["txtytz"].ravel(), [0.1, 0.2, 0.3])
    cv2 = pytest.importorskip("cv2")
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)  # the reader debug_code/showscan_node.cpp:193-202 uses
    np.testing.assert_array_equal(fs.getNode("extrinsicTlc").mat(), gtT)
    np.testing.assert_allclose(fs.getNode("RollPitchYaw").mat().ravel(), rpy)
    np.testing.assert_allclose(fs.getNode("txtytz").mat().ravel(), [0.1, 0.2, 0.3])


def test_keyframe_selection(oracle):
    from camlasercalibratool_b200 import formats as fmt

    base = fmt.CamPose(0.0, np.array([0, 0, 0, 1.0]), np.zeros(3))
    near = fmt.CamPose(1.0, np.array([0, 0, 0, 1.0]), np.array([0.05, 0, 0]))            # < 0.2 m, no rotation: dropped
    far = fmt.CamPose(2.0, np.array([0, 0, 0, 1.0]), np.array([0.25, 0, 0]))             # > 0.2 m: kept
    a = math.radians(12) / 2
    turned = fmt.CamPose(3.0, np.array([0, 0, math.sin(a), math.cos(a)]), np.array([0.25, 0, 0]))  # 12 deg: kept
    small = fmt.CamPose(4.0, np.array([0, 0, math.sin(a), math.cos(a)]), np.array([0.30, 0, 0]))   # 5 cm, 0 deg: dropped
    kept = fmt.select_keyframes([base, near, far, turned, small])
    assert [k.timestamp for k in kept] == [0.0, 2.0, 3.0]


@pytest.mark.gpu
def test_offline_pipeline_without_ros(tmp_path, oracle):
    """apriltag_pose.txt + laser segments -> key frames -> nearest pose -> batched line fits -> closed form -> LM ->
    result.yaml, against the same chain done with the oracle."""
    from camlasercalibratool_b200 import formats as fmt

    p = oracle.generate(60, 180, seed=5, sigma=0.005)
    poses = _poses_from_problem(p, fmt)
    fmt.save_cam_pose_txt(tmp_path / "apriltag_pose.txt", poses)
    tagpose = fmt.load_cam_pose_txt(tmp_path / "apriltag_pose.txt")
    scans = []
    for f in range(p.n_frames):
        pts = p.points[p.offsets[f]:p.offsets[f + 1]]
        scans.append((100.0 + 0.5 * f + 0.004, pts))            # 4 ms after the image: matched
        scans.append((100.0 + 0.5 * f + 0.2, pts + 0.5))        # 200 ms off: no pose within 20 ms, dropped
    Tlc, rep = fmt.calibrate_offline(tagpose, scans, result_yaml=str(tmp_path / "result.yaml"))
    assert rep["n_obs"] > 40
    # the same chain with the oracle (pose file precision: 1e-10)
    kept = {k.timestamp for k in fmt.select_keyframes(tagpose)}
    frames = [f for f in range(p.n_frames) if (100.0 + 0.5 * f) in kept and p.offsets[f + 1] > p.offsets[f]]
    fp = np.array([np.concatenate([fmt.quat_inverse(tagpose[f].qwc), -fmt.quat_to_rot(fmt.quat_inverse(tagpose[f].qwc)) @ tagpose[f].twc])
                   for f in frames])
    segs = [p.points[p.offsets[f]:p.offsets[f + 1]] for f in frames]
    ends = []
    for pts in segs:
        line, _, _ = oracle.line_fit(pts, (0.0, 0.0))
        x_s, x_e, y_s, y_e = pts[0, 0], pts[-1, 0], pts[0, 1], pts[-1, 1]
        if abs(x_e - x_s) > abs(y_e - y_s):
            y_s, y_e = -(x_s * line[0] + 1) / line[1], -(x_e * line[0] + 1) / line[1]
        else:
            x_s, x_e = -(y_s * line[1] + 1) / line[0], -(y_e * line[1] + 1) / line[0]
        ends.append(np.array([[x_s, y_s, 0.0], [x_e, y_e, 0.0]]))
    on_line = oracle.Problem(fp, np.arange(len(frames) + 1) * 2, np.concatenate(ends))
    Tlc0, _, _, _ = oracle.closed_form(on_line)
    np.testing.assert_allclose(rep["Tlc_closed_form"], Tlc0, atol=1e-7)
    full = oracle.Problem(fp, np.concatenate([[0], np.cumsum([len(s) for s in segs])]), np.concatenate(segs))
    xo, _, _ = oracle.solve(full, oracle.T_to_pose7(np.linalg.inv(rep["Tlc_closed_form"])))
    np.testing.assert_allclose(Tlc, np.linalg.inv(oracle.pose7_to_T(xo)), atol=1e-7)
    assert np.abs(Tlc - oracle.ground_truth()[0]).max() < 5e-3
    saved = fmt.read_result_yaml(str(tmp_path / "result.yaml"))
    np.testing.assert_array_equal(saved["extrinsicTlc"], Tlc)


def synthetic_scans(n_scans, n_beams=1081, seed=0):
    """LaserScan-like ranges: smooth walls at 3.5-6.5 m, 1 % dropped beams, and (3 of 4 scans) a flat board 0.6-1.5 m away
    inside the front sector."""
    rng = np.random.default_rng(seed)
    a0, inc = -2.356, 4.712 / (n_beams - 1)
    ang = a0 + np.arange(n_beams) * inc
    out = np.empty((n_scans, n_beams), dtype=np.float32)
    for k in range(n_scans):
        r = (5 + np.sin(ang * 3 + rng.uniform(0, 6)) * 1.5 + rng.normal(size=n_beams) * 0.01).astype(np.float32)
        if k % 4 != 0:
            c, w, d = rng.uniform(-0.6, 0.6), rng.uniform(0.15, 0.35), rng.uniform(0.6, 1.5)
            m = np.abs(ang - c) < w
            r[m] = (d / np.cos(ang[m] - c) + rng.normal(size=int(m.sum())) * 0.003).astype(np.float32)
        r[rng.random(n_beams) < 0.01] = np.inf
        out[k] = r
    return out, a0, inc


def test_scan_preparation_restatements_agree(harness, oracle, oracle_np):
    """TranScanToPoints + AutoGetLinePts: C oracle == numpy twin == the product's host/device code (host build)."""
    import ctypes as C

    from camlasercalibratool_b200 import formats as fmt

    ranges, a0, inc = synthetic_scans(120, seed=3)
    harness.L.harness_auto_get_line_pts.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_double, C.c_double, C.c_double,
                                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
    found = 0
    for r in ranges:
        p = oracle.scan_to_points(r, a0, inc, 0.05)
        np.testing.assert_allclose(p, oracle_np.scan_to_points(r, a0, inc, 0.05), atol=1e-12)
        np.testing.assert_allclose(p, fmt.scan_to_points(r, a0, inc, 0.05), atol=1e-12)
        ref = oracle.auto_get_line_pts(p)
        assert ref == oracle_np.auto_get_line_pts(p)
        s, e = C.c_int(), C.c_int()
        harness.L.harness_auto_get_line_pts(r.ctypes.data_as(C.POINTER(C.c_float)), len(r), a0, inc, 0.05, C.byref(s), C.byref(e))
        assert (ref is None and s.value == -1) or ref == (s.value, e.value)
        found += ref is not None
    assert 60 < found <= 90  # the scans with a board (3 of 4) are found, the board-less ones are not
    # edge cases: empty scan, all-invalid scan, short scan
    assert oracle.auto_get_line_pts(np.zeros((0, 3))) is None
    assert oracle.auto_get_line_pts(oracle.scan_to_points(np.full(500, np.inf, dtype=np.float32), a0, inc, 0.05)) is None
    assert oracle.auto_get_line_pts(oracle.scan_to_points(np.full(40, 1.0, dtype=np.float32), a0, inc, 0.05)) is None


@pytest.mark.gpu
def test_batched_scan_segments_match_oracle(oracle):
    from camlasercalibratool_b200 import formats as fmt

    ranges, a0, inc = synthetic_scans(2000, seed=4)
    s, e = fmt.auto_get_line_segments(ranges, a0, inc, 0.05)
    for k in range(0, 2000, 7):
        ref = oracle.auto_get_line_pts(oracle.scan_to_points(ranges[k], a0, inc, 0.05))
        assert (ref is None and s[k] == -1 and e[k] == -1) or ref == (int(s[k]), int(e[k])), k
    segs = fmt.segments_from_scans(np.arange(2000) * 0.025, ranges, a0, inc, 0.05)
    assert len(segs) == int(np.sum(s >= 0)) and all(len(p) == e[k] - s[k] + 1 for (t, p), k in zip(segs, np.nonzero(s >= 0)[0]))


@pytest.mark.gpu
def test_save_plane_points_files(tmp_path, oracle):
    """planar.txt / RoiPoints.txt / RoiPtOnLines.txt (reference src/LaseCamCalCeres.cpp:68-110): the plane of every frame
    is (Tctag^-1)^T (0,0,1,0) and the points are moved by Tcl; three significant digits."""
    from camlasercalibratool_b200 import Oberserve
    from camlasercalibratool_b200.formats import save_plane_points

    p = oracle.generate(12, 40, seed=4, sigma=0.01)
    obs = [Oberserve(p.frame_pose[f, :4].copy(), p.frame_pose[f, 4:].copy(), p.points[p.offsets[f]:p.offsets[f + 1]],
                     p.points[p.offsets[f]:p.offsets[f + 1]][[0, -1]]) for f in range(p.n_frames)]
    Tcl = np.linalg.inv(oracle.ground_truth()[0])
    save_plane_points(obs, Tcl, str(tmp_path) + "/")
    planar = np.loadtxt(tmp_path / "planar.txt")
    roi = np.loadtxt(tmp_path / "RoiPoints.txt")
    lines = np.loadtxt(tmp_path / "RoiPtOnLines.txt")
    assert planar.shape == (12, 5) and roi.shape == (p.n_points, 4) and lines.shape == (24, 4)
    assert np.array_equal(planar[:, 0], np.arange(12))
    for f in range(12):
        T = np.eye(4)
        T[:3, :3] = oracle.quat_to_rot(p.frame_pose[f, :4])
        T[:3, 3] = p.frame_pose[f, 4:]
        want = np.linalg.inv(T).T @ np.array([0, 0, 1.0, 0])
        np.testing.assert_allclose(planar[f, 1:], want, rtol=6e-3, atol=6e-3)  # %.3g
    cam = p.points @ Tcl[:3, :3].T + Tcl[:3, 3]
    np.testing.assert_allclose(roi[:, 1:], cam, rtol=6e-3, atol=1e-12)
    assert np.array_equal(roi[:, 0], np.repeat(np.arange(12), np.diff(p.offsets)))
    for tok in (tmp_path / "planar.txt").read_text().split():
        assert "%.3g" % float(tok) == tok  # std::setprecision(3), default float format


def test_offline_driver_refuses_too_little_data():
    """reference main/calibr_offline.cpp:55-59 (fewer than 10 tag poses) and :158-163 (fewer than 5 matched observations):
    both exits are taken before any numeric work, so they are testable without a GPU."""
    from camlasercalibratool_b200.formats import CamPose, calibrate_offline, observations_from_segments

    poses = [CamPose(0.1 * i, np.array([0, 0, 0, 1.0]), np.array([0.0, 0.0, 1.0 + 0.3 * i])) for i in range(9)]
    Tlc, why = calibrate_offline(poses, [])
    assert Tlc is None and why == "apriltag pose less than 10."
    poses.append(CamPose(0.9, np.array([0, 0, 0, 1.0]), np.array([0.0, 0.0, 4.0])))
    # scans whose time stamps are more than 20 ms away from every pose are not matched (:116)
    scans = [(0.1 * i + 0.05, np.array([[1.0, 0.1 * i, 0.0], [1.0, 0.1 * i + 0.2, 0.0]])) for i in range(10)]
    assert observations_from_segments(poses, scans) == []
    Tlc, why = calibrate_offline(poses, scans)
    assert Tlc is None and why == "Valid Calibra Data Less"
    # empty segments are skipped, too
    assert observations_from_segments(poses, [(0.1, np.zeros((0, 3)))]) == []


def test_scan_segmentation_soak_with_pathological_scans(harness, oracle, oracle_np):
    """600 random scans -- 0 to 2000 beams, boards of any width and distance, NaN / inf / zero / negative / huge ranges,
    constant scans -- through the C oracle, the literal numpy twin and the host build of the device code: always the same
    answer (a 3000-scan run of the same loop was clean)."""
    import ctypes as C

    L = harness.L
    L.harness_auto_get_line_pts.argtypes = [C.POINTER(C.c_float), C.c_int64, C.c_double, C.c_double, C.c_double,
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(5)
    found = 0
    for _ in range(600):
        n = int(rng.choice([0, 1, 5, 40, 99, 100, 101, 360, 720, 1081, 2000]))
        a0, inc = rng.uniform(-3.2, 0), rng.uniform(0.001, 0.02)
        ang = a0 + np.arange(n) * inc
        r = (rng.uniform(1, 8) + np.sin(ang * rng.uniform(1, 5) + rng.uniform(0, 6)) * rng.uniform(0, 2)
             + rng.normal(size=n) * 0.01).astype(np.float32)
        kind = int(rng.integers(0, 8))
        if n > 10 and kind < 5:
            c, w, d = rng.uniform(ang[0], ang[-1]), rng.uniform(0.02, 0.6), rng.uniform(0.2, 3.0)
            m = np.abs(ang - c) < w
            r[m] = (d / np.cos(np.clip(ang[m] - c, -1.4, 1.4)) + rng.normal(size=int(m.sum())) * rng.choice([0, 0.003, 0.02])).astype(np.float32)
        if kind == 5 and n:
            r[rng.random(n) < 0.3] = np.nan
        if kind == 6 and n:
            r[rng.random(n) < 0.3] = rng.choice([0.0, -1.0, np.inf, 1e9])
        if kind == 7 and n:
            r[:] = rng.choice([0.5, 2.0, 40.0])
        rmin = float(rng.choice([0.05, 0.0, 0.5]))
        ref = oracle.auto_get_line_pts(oracle.scan_to_points(r, a0, inc, rmin)) if n else None
        twin = oracle_np.auto_get_line_pts(oracle_np.scan_to_points(r, a0, inc, rmin)) if n else None
        s, e = C.c_int(), C.c_int()
        L.harness_auto_get_line_pts(r.ctypes.data_as(C.POINTER(C.c_float)), n, a0, inc, rmin, C.byref(s), C.byref(e))
        got = None if s.value < 0 else (s.value, e.value)
        assert ref == twin == got
        found += ref is not None
    assert found > 10

"""Camera measurement chain of the synthetic generator: camodocal pinhole-radtan / equidistant
models (camera_models/src/PinholeCamera.cc, EquidistantCamera.cc), kalibr grid corners (src/calcCamPose.cpp:107-136) and
the planar PnP the reference gets from cv::solvePnP (src/calcCamPose.cpp:225).
CPU: the product's host/device code (host build) vs the literal numpy twin (np.roots for the equidistant back-projection)
and vs OpenCV's solvePnP itself.  GPU: the device generator with a camera model vs the same chain on the host."""
import ctypes as C

import numpy as np
import pytest

PINHOLE = np.array([367.05, 366.94, 368.72, 241.14, -0.28, 0.07, 0.0003, -0.0002])       # fx fy cx cy k1 k2 p1 p2
PINHOLE_NODIST = np.array([367.05, 366.94, 368.72, 241.14, 0.0, 0.0, 0.0, 0.0])
EQUI = np.array([363.0, 363.2, 370.1, 240.3, -0.013, 0.021, -0.034, 0.012])              # mu mv u0 v0 k2 k3 k4 k5
GRID = (6, 6, 0.055, 0.3)


def _bind(h):
    dp = C.POINTER(C.c_double)
    L = h.L
    L.harness_camera_project.argtypes = [C.c_int, dp, dp, dp]
    L.harness_camera_lift.argtypes = [C.c_int, dp, dp, dp]
    L.harness_grid_corners.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, dp]
    L.harness_pixel_noise.argtypes = [C.c_uint64, C.c_double, C.c_int64, C.c_int, dp]
    L.harness_pnp_planar.argtypes = [C.c_int, dp, dp, dp, dp]
    L.harness_camera_estimate_pose.argtypes = [C.c_int, dp, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint64,
                                               C.c_int64, dp, dp, C.POINTER(C.c_float)]
    return L


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def random_board_pose(rng):
    """A board 0.5-2 m in front of the camera, tilted by up to ~35 degrees, roughly centred."""
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = rng.uniform(0, 0.6)
    q = np.concatenate([ax * np.sin(ang / 2), [np.cos(ang / 2)]])
    t = np.array([rng.uniform(-0.4, 0.1), rng.uniform(-0.35, 0.05), rng.uniform(0.5, 2.0)])
    return np.concatenate([q, t])


@pytest.mark.parametrize("model,k", [(1, PINHOLE), (1, PINHOLE_NODIST), (2, EQUI), (2, np.array([363.0, 363.2, 370.1, 240.3, 0, 0, 0, 0.0]))])
def test_camera_models_match_the_literal_restatement(harness, oracle_np, model, k):
    L = _bind(harness)
    rng = np.random.default_rng(0)
    for _ in range(300):
        P = np.array([rng.uniform(-1, 1), rng.uniform(-0.8, 0.8), rng.uniform(0.4, 3.0)])
        uv = np.empty(2)
        L.harness_camera_project(model, _d(k), _d(P), _d(uv))
        np.testing.assert_allclose(uv, oracle_np.camera_project(model, k, P), rtol=0, atol=1e-9)
        xy = np.empty(2)
        L.harness_camera_lift(model, _d(k), _d(uv), _d(xy))
        np.testing.assert_allclose(xy, oracle_np.camera_lift_normalised(model, k, uv), rtol=0, atol=2e-9)
        # and the lift undoes the projection (to the accuracy of the 8-step recursion for radtan)
        # (radtan: the reference's 8-step recursion has not converged far off-axis -- reproduced, not "fixed")
        if model == 2 or np.hypot(*(P[:2] / P[2])) < 0.45:
            np.testing.assert_allclose(xy, P[:2] / P[2], atol=2e-5 if model == 1 else 1e-9)


def test_grid_corners_and_pixel_noise(harness, oracle_np):
    L = _bind(harness)
    xy = np.empty((144, 2))
    L.harness_grid_corners(6, 6, 0.055, 0.3, _d(xy))
    np.testing.assert_allclose(xy, oracle_np.grid_corners(*GRID), atol=1e-15)
    assert np.allclose(xy[:4], [[0, 0], [0.055, 0], [0.055, 0.055], [0, 0.055]]) and np.isclose(xy[4, 0], 0.0715)
    n = np.array([np.empty(2) for _ in range(4000)])
    for i in range(4000):
        L.harness_pixel_noise(7, 0.25, i // 144, i % 144, _d(n[i]))
    assert abs(n.mean()) < 0.02 and abs(n.std() - 0.25) < 0.01
    z = np.empty(2)
    L.harness_pixel_noise(7, 0.0, 3, 5, _d(z))
    assert not z.any()


def test_planar_pnp_matches_opencv(harness, oracle_np):
    """The product's planar PnP vs cv2.solvePnP (the reference's dependency) on identical float32 inputs."""
    cv2 = pytest.importorskip("cv2")
    L = _bind(harness)
    rng = np.random.default_rng(1)
    corners = oracle_np.grid_corners(*GRID)
    for trial in range(40):
        fp = random_board_pose(rng)
        noise = rng.normal(size=(144, 2)) * (0.0 if trial % 2 == 0 else 0.3)
        Rcv, tcv, p2 = oracle_np.estimate_pose_cv(1, PINHOLE, corners, fp, noise)
        obj = corners.astype(np.float32).astype(np.float64)
        img = p2.astype(np.float64)
        R, t = np.empty(9), np.empty(3)
        assert L.harness_pnp_planar(144, _d(np.ascontiguousarray(obj)), _d(np.ascontiguousarray(img)), _d(R), _d(t)) == 1
        R = R.reshape(3, 3)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
        # same minimum of the reprojection error as OpenCV (which stops at FLT_EPSILON after <= 20 iterations)
        ang = np.arccos(np.clip((np.trace(R.T @ Rcv) - 1) / 2, -1, 1))
        assert ang < 2e-5 and np.linalg.norm(t - tcv) < 2e-5, (trial, ang, np.linalg.norm(t - tcv))

        def cost(Rm, tm):
            pc = np.c_[obj, np.zeros(144)] @ Rm.T + tm
            return np.sum((pc[:, :2] / pc[:, 2:] - img) ** 2)

        assert cost(R, t) <= cost(Rcv, tcv) * (1 + 1e-6) + 1e-18  # never worse than OpenCV's answer
        if trial % 2 == 0:  # noise-free: the true pose up to the float32 quantisation of the points
            Rt = oracle_np.quat_to_rot(fp[:4])
            assert np.abs(R - Rt).max() < 1e-4 and np.abs(t - fp[4:]).max() < 1e-4


@pytest.mark.parametrize("model,k", [(1, PINHOLE), (2, EQUI)])
def test_full_chain_matches_opencv_chain(harness, oracle_np, model, k):
    cv2 = pytest.importorskip("cv2")
    L = _bind(harness)
    rng = np.random.default_rng(2)
    corners = oracle_np.grid_corners(*GRID)
    for frame in range(12):
        fp = random_board_pose(rng)
        est = np.empty(7)
        uv = np.empty(288, dtype=np.float32)
        ok = L.harness_camera_estimate_pose(model, _d(k), 0.2, 6, 6, 0.055, 0.3, 99, frame, _d(fp), _d(est),
                                            uv.ctypes.data_as(C.POINTER(C.c_float)))
        assert ok == 1
        noise = np.empty((144, 2))
        for i in range(144):
            L.harness_pixel_noise(99, 0.2, frame, i, _d(noise[i]))
        Rcv, tcv, p2 = oracle_np.estimate_pose_cv(model, k, corners, fp, noise)
        np.testing.assert_allclose(uv.reshape(144, 2), p2, rtol=0, atol=3e-7)  # float32 normalised points
        Re = oracle_np.quat_to_rot(est[:4])
        ang = np.arccos(np.clip((np.trace(Re.T @ Rcv) - 1) / 2, -1, 1))
        assert ang < 3e-5 and np.linalg.norm(est[4:] - tcv) < 3e-5
        # pose noise of the expected size: 0.2 px at f = 365 px over a 0.4 m board 0.5-2 m away -> well below a degree
        Rt = oracle_np.quat_to_rot(fp[:4])
        assert np.arccos(np.clip((np.trace(Re.T @ Rt) - 1) / 2, -1, 1)) < 0.02 and np.linalg.norm(est[4:] - fp[4:]) < 0.02


def test_camera_mode_pose_draw_keeps_the_grid_in_the_image(harness, oracle_np):
    """Generator with a camera model: the drawn board pose satisfies the plain generator's rules AND every grid corner
    projects into the image (the tag detector of the reference needs the whole grid, kalibratag_detector_node.cpp)."""
    L = _bind(harness)
    dp = C.POINTER(C.c_double)
    L.harness_gen_frame_pose_camera.argtypes = [C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                                C.c_uint64, C.c_int64, C.c_int, dp]
    corners = oracle_np.grid_corners(*GRID)
    for model, k in [(1, PINHOLE), (2, EQUI)]:
        n_ok = 0
        for frame in range(200):
            fp = np.empty(7)
            ok = L.harness_gen_frame_pose_camera(model, _d(k), 6, 6, 0.055, 0.3, 752, 480, 5, frame, 1, _d(fp))
            n_ok += ok
            if ok:
                R = oracle_np.quat_to_rot(fp[:4])
                P = corners[:, :2] @ R[:, :2].T + fp[4:]
                assert (P[:, 2] > 0.05).all()
                uv = np.array([oracle_np.camera_project(model, k, p) for p in P])
                assert (uv[:, 0] >= 0).all() and (uv[:, 0] < 752).all() and (uv[:, 1] >= 0).all() and (uv[:, 1] < 480).all()
        assert n_ok == 200  # 512 attempts are plenty for the generator's pose distribution


# ---- GPU: the device generator with a camera model ---------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("camera,model", [("radtan", 1), ("equi", 2)])
def test_device_generator_camera_chain_matches_the_host_chain(harness, oracle, camera, model):
    from camlasercalibratool_b200 import Problem
    from camlasercalibratool_b200.api import CAMERA_DEFAULTS

    L = _bind(harness)
    dp = C.POINTER(C.c_double)
    L.harness_gen_frame_pose_camera.argtypes = [C.c_int, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                                C.c_uint64, C.c_int64, C.c_int, dp]
    k = np.array(CAMERA_DEFAULTS[camera])
    N, M, seed, px = 64, 120, 31, 0.3
    with Problem.synthetic(N, M, seed=seed, sigma=0.01, with_edges=True, camera=camera, pixel_sigma=px) as g:
        d = g.download()
        true = g.download_true_poses()
        for i in range(N):
            fp, est = np.empty(7), np.empty(7)
            uv = np.empty(288, dtype=np.float32)
            assert L.harness_gen_frame_pose_camera(model, _d(k), 6, 6, 0.055, 0.3, 752, 480, seed, i, 1, _d(fp)) == 1
            np.testing.assert_allclose(true[i], fp, rtol=0, atol=1e-13)
            assert L.harness_camera_estimate_pose(model, _d(k), px, 6, 6, 0.055, 0.3, seed, i, _d(fp), _d(est),
                                                  uv.ctypes.data_as(C.POINTER(C.c_float))) == 1
            # device libm vs host libm differ in the last bits; a float32 rounding of a normalised point may flip (6e-8),
            # which moves the PnP estimate by ~1e-7
            np.testing.assert_allclose(d["frame_pose"][i], est, rtol=0, atol=2e-6)
        # pose noise of the expected size, and not zero
        dev = np.abs(d["frame_pose"] - true).max(axis=0)
        assert 1e-6 < dev[4:].max() < 0.2 and dev[:4].max() < 0.1  # boards up to 5 m away: 0.3 px is several cm / a few degrees
        # the laser points lie on the TRUE board (sigma 0.01), not on the estimated one
        gtT, gt = oracle.ground_truth()
        ptrue = oracle.Problem(true, d["offsets"], d["points"], d["edge_points"])
        pest = oracle.Problem(d["frame_pose"], d["offsets"], d["points"], d["edge_points"])
        c_true = oracle.evaluate_normal(ptrue, gt)[0]
        c_est = oracle.evaluate_normal(pest, gt)[0]
        assert c_true < c_est
        # the solve on the device data = the oracle's solve on the downloaded data, and lands near the ground truth
        x0 = oracle.pose_plus(gt, np.array([0.02, -0.01, 0.03, 0.01, -0.02, 0.015]))
        xs, summ, _ = g.solve(x0)
        xo, so, _ = oracle.solve(pest, x0)
        ang, dt = oracle.pose_error(xs, xo)
        assert ang < 1e-6 and dt < 1e-6
        ang, dt = oracle.pose_error(xs, gt)
        assert ang < 0.05 and dt < 0.05


@pytest.mark.gpu
def test_device_generator_camera_without_pixel_noise_recovers_the_true_poses(oracle):
    from camlasercalibratool_b200 import Problem

    with Problem.synthetic(200, 64, seed=3, sigma=0.0, camera="equi", pixel_sigma=0.0) as g:
        d = g.download()
        true = g.download_true_poses()
        assert np.abs(d["frame_pose"] - true).max() < 2e-4  # float32 image/object points (cv::Point2f / Point3f)
        gt = oracle.ground_truth()[1]
        xs, summ, _ = g.solve(oracle.pose_plus(gt, np.array([0.02, -0.01, 0.03, 0.01, -0.02, 0.015])))
        ang, dt = oracle.pose_error(xs, gt)
        assert ang < 1e-3 and dt < 1e-3
    # without a camera the true poses are the poses
    with Problem.synthetic(20, 16, seed=3) as g:
        assert np.array_equal(g.download()["frame_pose"], g.download_true_poses())


# ---- board poses from tag detections (calcCamPose minus the detector) ---------------------------------------------------
def _synthetic_detections(oracle_np, model, k, rng, n_frames, sigma=0.15):
    """Per frame: a random subset of the 36 tags (ascending id), their 4 corners projected and disturbed, as float32."""
    corners = oracle_np.grid_corners(*GRID)[:, :2].reshape(36, 4, 2)
    out, truth = [], []
    for _ in range(n_frames):
        fp = random_board_pose(rng)
        R, t = oracle_np.quat_to_rot(fp[:4]), fp[4:]
        ids = np.sort(rng.choice(36, size=rng.integers(1, 37), replace=False)).astype(np.int32)
        uv = np.array([[oracle_np.camera_project(model, k, R @ np.array([X, Y, 0.0]) + t) for X, Y in corners[i]] for i in ids])
        out.append((ids, (uv + rng.normal(scale=sigma, size=uv.shape)).astype(np.float32)))
        truth.append(fp)
    return out, truth, corners


def _pose_cv(oracle_np, cv2, model, k, ids, uv, corners):
    p2 = np.array([oracle_np.camera_lift_normalised(model, k, p.astype(float)) for p in uv.reshape(-1, 2)], dtype=np.float32)
    p3 = np.c_[corners[ids].reshape(-1, 2), np.zeros(4 * len(ids))].astype(np.float32)
    _, rvec, tvec = cv2.solvePnP(p3, p2, np.eye(3, dtype=np.float32), np.zeros((1, 5), dtype=np.float32))
    Rcw, _ = cv2.Rodrigues(rvec)
    return Rcw.T, -Rcw.T @ tvec.ravel()  # T_wc (src/calcCamPose.cpp:229-230)


@pytest.mark.parametrize("model,k", [(1, PINHOLE), (2, EQUI)])
def test_pose_from_detections_matches_opencv(harness, oracle_np, model, k):
    cv2 = pytest.importorskip("cv2")
    L = _bind(harness)
    L.harness_estimate_pose_from_detections.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double,
                                                        C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    rng = np.random.default_rng(8)
    dets, truth, corners = _synthetic_detections(oracle_np, model, k, rng, 25)
    for (ids, uv), fp in zip(dets, truth):
        pose = np.empty(7)
        ok = L.harness_estimate_pose_from_detections(model, _d(k), 6, 6, 0.055, 0.3, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     np.ascontiguousarray(uv).ctypes.data_as(C.POINTER(C.c_float)), _d(pose))
        assert ok == 1
        Rwc, twc = _pose_cv(oracle_np, cv2, model, k, ids, uv, corners)
        Re = oracle_np.quat_to_rot(pose[:4])
        ang = np.arccos(np.clip((np.trace(Re.T @ Rwc) - 1) / 2, -1, 1))
        # a single tag (4 points, 5.5 cm) is poorly conditioned: OpenCV's own iteration stops at ~1e-4 there
        tol = 3e-5 if len(ids) >= 4 else 2e-3
        assert ang < tol and np.linalg.norm(pose[4:] - twc) < tol * 5, (len(ids), ang)
    # error convention: fewer than 4 points, or an id outside the grid -> false and the identity pose
    pose = np.empty(7)
    ids = np.array([40], dtype=np.int32)
    uv = np.zeros(8, dtype=np.float32)
    for n_det in (0, 1):
        assert L.harness_estimate_pose_from_detections(model, _d(k), 6, 6, 0.055, 0.3, n_det, ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       uv.ctypes.data_as(C.POINTER(C.c_float)), _d(pose)) == 0
        assert np.array_equal(pose, [0, 0, 0, 1, 0, 0, 0])


@pytest.mark.gpu
@pytest.mark.parametrize("camera,model,k", [("radtan", 1, PINHOLE), ("equi", 2, EQUI)])
def test_batched_board_poses_match_the_host_build(harness, oracle_np, tmp_path, camera, model, k):
    from camlasercalibratool_b200.formats import cam_poses_from_detections, estimate_board_poses, load_cam_pose_txt, save_cam_pose_txt

    L = _bind(harness)
    L.harness_estimate_pose_from_detections.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.c_double, C.c_double,
                                                        C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double)]
    rng = np.random.default_rng(11)
    dets, truth, _ = _synthetic_detections(oracle_np, model, k, rng, 300)
    dets[7] = (np.zeros(0, dtype=np.int32), np.zeros((0, 4, 2), dtype=np.float32))   # nothing detected
    dets[9] = (np.array([36], dtype=np.int32), np.zeros((1, 4, 2), dtype=np.float32))  # id outside the 6 x 6 grid
    pose, ok = estimate_board_poses(camera, dets, intrinsics=k)
    assert not ok[7] and not ok[9] and ok.sum() == 298
    assert np.array_equal(pose[7], [0, 0, 0, 1, 0, 0, 0])
    for i, (ids, uv) in enumerate(dets):
        want = np.empty(7)
        good = L.harness_estimate_pose_from_detections(model, _d(k), 6, 6, 0.055, 0.3, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       np.ascontiguousarray(uv).ctypes.data_as(C.POINTER(C.c_float)), _d(want))
        assert bool(good) == bool(ok[i])
        tol = 2e-6 if len(ids) >= 4 else 1e-4
        np.testing.assert_allclose(pose[i], want, rtol=0, atol=tol)
        if ok[i] and len(ids) >= 9:  # T_wc is the inverse of the board pose the pixels were made from
            Rca = oracle_np.quat_to_rot(truth[i][:4])
            assert np.abs(oracle_np.quat_to_rot(pose[i, :4]) - Rca.T).max() < 0.05
            assert np.abs(pose[i, 4:] + Rca.T @ truth[i][4:]).max() < 0.05
    # and on to apriltag_pose.txt, as the detector node does
    cams = cam_poses_from_detections(np.arange(300) * 0.05, camera, dets, intrinsics=k)
    assert len(cams) == 298
    save_cam_pose_txt(tmp_path / "apriltag_pose.txt", cams)
    back = load_cam_pose_txt(tmp_path / "apriltag_pose.txt")
    assert len(back) == 298 and abs(back[8].timestamp - 0.5) < 1e-9  # frames 7 and 9 were dropped

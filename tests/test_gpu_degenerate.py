"""GPU parity of the SOLVE on the degenerate board geometries the reference teaches with (-m gpu).

reference main/calibr_simulation.cpp:39-56 invites the user to sample the board poses about fewer camera axes ("Try me!!!
ONLY pitch, Wow!!!! you will find we can not estimate the tlc.z()") and src/LaseCamCalCeres.cpp:364-379 then prints the
null space of H.  On such data the normal equations are rank deficient: the damped 6x6 Cholesky step of the device LM
(csrc/clc_lm.cuh) and the Householder QR of [J; D] that Ceres' DENSE_QR (and the oracle) uses see different condition
numbers.  What must agree is everything the data determine: the cost reached, the pose projected on the observable
directions, and the answer the reference's analysis tail gives (which directions are unobservable).  Along a null-space
direction the data say nothing and the two arithmetic paths may legitimately stop at different points.

The generator below restates calibr_simulation.cpp:34-103 with a seeded RNG and the reference's commented variants.
Nothing here reads /root/reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
RLC = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])  # calibr_simulation.cpp:15-19
TLC = np.array([0.1, 0.2, 0.3])


def rot(axis, a):
    c, s = np.cos(a), np.sin(a)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def simulate(oracle, variant, n_frames=50, beams=180, seed=0, sigma=0.0, centred=False):
    """calibr_simulation.cpp:34-103; variant in {"all", "no_yaw", "only_pitch", "only_roll"} (:41-56); centred = the
    commented `tca(0, 0, z)` of :59."""
    rng = np.random.default_rng(seed)
    fp, pts, off = [], [], [0]
    for _ in range(n_frames):
        r, p, y = rng.uniform(-np.pi / 6, np.pi / 6, size=3)
        if variant == "all":
            Rca = rot("z", y) @ rot("y", p) @ rot("x", r)
        elif variant == "no_yaw":
            Rca = rot("y", p) @ rot("x", r)
        elif variant == "only_pitch":
            Rca = rot("y", p)
        else:
            Rca = rot("x", r)
        tca = np.array([0.0, 0.0, rng.uniform(1, 5)]) if centred else np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(1, 5)])
        Tla = np.eye(4)
        Tla[:3, :3] = RLC @ Rca
        Tla[:3, 3] = RLC @ tca + TLC
        plane = np.linalg.inv(Tla).T @ np.array([0.0, 0.0, 1.0, 0.0])
        n, d = plane[:3], plane[3]
        frame = []
        for j in range(beams):
            th = -np.pi / 2 + j * np.pi / beams
            ray = np.array([np.cos(th), np.sin(th), 0.0])
            with np.errstate(divide="ignore", invalid="ignore"):
                depth = -d / (ray @ n)
            if not np.isfinite(depth) or depth < 0:
                continue
            q = (depth + (rng.normal() * sigma if sigma > 0 else 0.0)) * ray
            if abs(q[0]) < 5 and abs(q[1]) < 5:
                frame.append(q)
        fp.append(np.concatenate([oracle.rot_to_quat(Rca), tca]))
        pts.extend(frame)
        off.append(off[-1] + len(frame))
    return oracle.Problem(np.array(fp), np.array(off, dtype=np.int64), np.array(pts).reshape(-1, 3))


def gpu_problem(p):
    from camlasercalibratool_b200 import Problem

    return Problem.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points, use_loss=p.use_loss, cauchy_a=p.cauchy_a)


def local_difference(oracle, xa, xb):
    """6-vector (dt, dtheta) of xb relative to xa in the parameterisation of PoseLocalParameterization::Plus."""
    Ta, Tb = np.asarray(oracle.pose7_to_T(xa)).reshape(4, 4), np.asarray(oracle.pose7_to_T(xb)).reshape(4, 4)
    dR = Ta[:3, :3].T @ Tb[:3, :3]
    w = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    return np.concatenate([Tb[:3, 3] - Ta[:3, 3], w])


@pytest.mark.parametrize("variant,centred,expect_null", [
    ("all", False, 0), ("no_yaw", False, 0), ("only_pitch", False, None), ("only_roll", False, None),
    ("only_pitch", True, None), ("only_roll", True, None)])
@pytest.mark.parametrize("sigma", [0.0, 0.01])
def test_solve_on_the_reference_teaching_geometries(oracle, variant, centred, expect_null, sigma):
    p = simulate(oracle, variant, seed=11, sigma=sigma, centred=centred)
    gt = oracle.ground_truth()[1]
    starts = [X0, oracle.pose_plus(gt, np.array([0.05, -0.04, 0.03, 0.02, -0.03, 0.025]))]
    with gpu_problem(p) as g:
        for x0 in starts:
            x, s, tr = g.solve(x0)
            xo, so, tro = oracle.solve(p, x0)
            # the reference's own diagnostic at the solution: which directions does H not see (:364-379)
            H, b, chi, sv = g.information(x)
            V = g.last_V
            rH, rb, rchi, rsv = oracle.information(p, xo)
            n_null = int(np.sum(sv < 1e-8 * max(sv[0], 1e-300)))
            if expect_null is not None:
                assert n_null == expect_null, (variant, sv)
            ctx = f"{variant} centred={centred} sigma={sigma} null={n_null} term {s.termination}/{so.termination} it {s.num_iterations}/{so.num_iterations}"
            if n_null == 0:
                # well posed (the reference's remark: yaw does not matter): the north-star tolerance applies in full
                ang, dt = oracle.pose_error(x, xo)
                assert ang < 1e-6 and dt < 1e-6, (ctx, ang, dt)
                assert s.termination == so.termination and s.num_iterations == so.num_iterations, ctx
                np.testing.assert_allclose(sv, rsv, rtol=1e-6)
            else:
                # rank deficient: compare what the data determine
                obs_dirs = V[:, : 6 - n_null]  # observable subspace at the GPU solution
                diff = local_difference(oracle, x, xo)
                assert np.abs(obs_dirs.T @ diff).max() < 1e-6, (ctx, obs_dirs.T @ diff)
                # both minimisers reached the same cost level (absolute 1e-12 covers the noise-free cost ~ 0)
                assert abs(s.final_cost - so.final_cost) <= 1e-6 * so.final_cost + 1e-12, (ctx, s.final_cost, so.final_cost)
                assert s.termination != 6 and so.termination != 6, ctx  # neither path FAILS (CLC_TERM_FAILURE)
                # the same number of unobservable directions is reported for the oracle's solution
                assert int(np.sum(rsv < 1e-8 * max(rsv[0], 1e-300))) == n_null, (ctx, rsv)
            if sigma == 0.0 and n_null == 0:
                ang, dt = oracle.pose_error(x, gt)
                assert ang < 1e-9 and dt < 1e-9, (ctx, ang, dt)


def test_unobservable_translation_is_the_one_the_reference_names(oracle):
    """"ONLY pitch ... we can not estimate the tlc.z()" (calibr_simulation.cpp:50): with boards rotated about the camera's
    y axis only, every board normal lies in the camera x-z plane, so a translation of T_cl along camera y changes no
    residual: the null space printed by the analysis tail must contain exactly that direction (t_lc.z = -(R_lc t_cl).z,
    and R_lc maps camera y to laser -z)."""
    p = simulate(oracle, "only_pitch", seed=3, sigma=0.0)
    gt = oracle.ground_truth()[1]
    with gpu_problem(p) as g:
        H, b, chi, sv = g.information(gt)
        V = g.last_V
    null = V[:, sv < 1e-8 * sv[0]]
    assert null.shape[1] >= 1
    e_ty = np.zeros(6)
    e_ty[1] = 1.0
    # e_ty lies in the span of the null-space columns
    resid = e_ty - null @ (null.T @ e_ty)
    assert np.linalg.norm(resid) < 1e-6, (sv, null)

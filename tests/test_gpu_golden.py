"""GPU parity against the COMMITTED golden fixtures (tests/golden/, produced by tests/golden/make_golden.py from sympy and
the numpy/LAPACK twin): the CUDA path through the C ABI, without the C oracle in between (it only re-creates the seeded
input arrays the fixtures were computed on, and those are checked against the fixture's own checksums first)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
TERM = {1: "CONVERGENCE_FUNCTION", 2: "CONVERGENCE_PARAMETER", 3: "CONVERGENCE_GRADIENT"}


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def frame_with_plane(n, d):
    """A board pose (qx qy qz qw tx ty tz) whose plane in the camera frame is n.x + d = 0: third rotation column n, t = -d n."""
    n = np.asarray(n, dtype=float)
    a = np.cross(n, [1.0, 0, 0] if abs(n[0]) < 0.9 else [0, 1.0, 0])
    a /= np.linalg.norm(a)
    R = np.stack([a, np.cross(n, a), n], axis=1)
    from camlasercalibratool_b200.api import T_to_pose7

    T = np.eye(4)
    T[:3, :3] = R
    q = T_to_pose7(T)[3:]
    return np.concatenate([q, -d * n])


def test_factor_known_answers_through_the_sweep_kernel():
    """sympy-derived (r, J) of PointInPlaneFactor (reference src/LaseCamCalCeres.cpp:43-66).  The factor's scale 1/sqrt(M) is
    realised by a frame of M = 1/scale^2 identical points: cost = M r^2 / 2, H = M J^T J, g = M J^T r (loss off)."""
    from camlasercalibratool_b200 import Problem

    for case in load("factor_kat.json"):
        M = int(round(1.0 / case["scale"] ** 2))
        fp = frame_with_plane(case["plane"][:3], case["plane"][3])[None, :]
        pts = np.tile(case["pt"], (M, 1))
        r, J = case["r"], np.array(case["J"])
        with Problem.from_arrays(fp, np.array([0, M]), pts, use_loss=False) as g:
            np.testing.assert_allclose(g.download()["planes"][0], case["plane"], atol=1e-15)
            cost, H, grad = g.eval(case["pose7"])
        assert abs(cost - 0.5 * M * r * r) <= 1e-14 * max(1.0, M * r * r)
        np.testing.assert_allclose(H, M * np.outer(J, J), rtol=0, atol=1e-13 * max(1.0, M * (J @ J)))
        np.testing.assert_allclose(grad, M * J * r, rtol=0, atol=1e-13 * max(1.0, M * abs(r) * np.abs(J).max()))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_config1_against_the_fixtures(oracle, seed):
    """BASELINE config 1 (50 x 180, the reference's own size): sums, LM trajectory, closed form and analysis tail."""
    from camlasercalibratool_b200 import Problem

    gold = load(f"config1_seed{seed}.json")
    p = oracle.generate(50, 180, seed=seed, sigma=0.01)  # the seeded inputs; pinned by the fixture's own checksums:
    assert p.n_points == gold["n_points"] and p.offsets.tolist() == gold["offsets"]
    checksum = float(np.sum(p.points * np.arange(1, 3 * p.n_points + 1).reshape(-1, 3) % 7))
    assert abs(checksum - gold["points_checksum"]) <= 1e-12 * abs(gold["points_checksum"])
    with Problem.from_arrays(p.frame_pose, p.offsets, p.points) as g:
        for name in ("identity", "ground_truth"):
            e = gold["eval_" + name]
            cost, H, grad = g.eval(e["pose7"])
            scale = np.abs(e["H"]).max()
            assert abs(cost - e["cost"]) <= 1e-12 * abs(e["cost"])
            np.testing.assert_allclose(H, e["H"], rtol=0, atol=1e-11 * scale)
            np.testing.assert_allclose(grad, e["g"], rtol=0, atol=1e-11 * scale)
        x, s, tr = g.solve(X0)
        assert TERM[s.termination] == gold["solve"]["termination"]
        ang, dt = oracle.pose_error(x, gold["solve"]["pose7"])
        assert ang < 1e-6 and dt < 1e-6  # north_star tolerance; in practice ~1e-10
        n = len(gold["solve"]["costs"])
        np.testing.assert_allclose([t.cost for t in tr][:n], gold["solve"]["costs"], rtol=1e-8)
        assert [bool(t.step_is_successful) for t in tr][:n] == gold["solve"]["accepted"]
        np.testing.assert_allclose([t.trust_region_radius for t in tr][:n], gold["solve"]["radius"], rtol=1e-6)
        T, un, AtA, Atb = g.closed_form()
        np.testing.assert_allclose(T, gold["closed_form"]["Tlc"], atol=1e-8)
        assert un == gold["closed_form"]["unobservable"]
        np.testing.assert_allclose(Atb, gold["closed_form"]["Atb"], rtol=1e-11, atol=1e-11)
        assert abs(np.trace(AtA) - gold["closed_form"]["AtA_trace"]) <= 1e-11 * gold["closed_form"]["AtA_trace"]
        H, b, chi, sv = g.information(gold["solve"]["pose7"])
        np.testing.assert_allclose(H, gold["information"]["H"], atol=1e-9)
        np.testing.assert_allclose(b, gold["information"]["b"], atol=1e-9)
        assert abs(chi - gold["information"]["chi"]) < 1e-11
        np.testing.assert_allclose(sv, gold["information"]["singular_values"], rtol=1e-8)

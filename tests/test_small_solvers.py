"""Property tests (hypothesis) of the small dense pieces the device code is built from -- host build of the same source:
Eigen's quaternion <-> rotation conversions (reference src/LaseCamCalCeres.cpp:215-219, :311-314), the 6x6 Cholesky step of
the LM update, the pivoted elimination of the PnP, and the Newton inverse of the Kannala-Brandt polynomial."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

dp = C.POINTER(C.c_double)
finite = st.floats(min_value=-1.0, max_value=1.0, allow_nan=False, allow_infinity=False)


def _d(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def L(harness):
    lib = harness.L
    lib.harness_quat_to_rot.argtypes = [dp, dp]
    lib.harness_rot_to_quat.argtypes = [dp, dp]
    lib.harness_chol6_solve.argtypes = [dp, dp, dp]
    lib.harness_solve_linear.argtypes = [dp, dp, C.c_int]
    lib.harness_equi_r.argtypes = [dp, C.c_double]
    lib.harness_equi_r.restype = C.c_double
    lib.harness_equi_theta_from_r.argtypes = [dp, C.c_double]
    lib.harness_equi_theta_from_r.restype = C.c_double
    return lib


@settings(max_examples=300, deadline=None)
@given(st.tuples(finite, finite, finite, finite))
def test_quaternion_rotation_round_trip(L, q):
    q = np.array(q)
    n = np.linalg.norm(q)
    if n < 1e-3:
        return
    q /= n
    R, q2 = np.empty(9), np.empty(4)
    L.harness_quat_to_rot(_d(q), _d(R))
    Rm = R.reshape(3, 3)
    np.testing.assert_allclose(Rm @ Rm.T, np.eye(3), atol=1e-14)
    assert abs(np.linalg.det(Rm) - 1) < 1e-13
    L.harness_rot_to_quat(_d(R), _d(q2))
    # same rotation (q and -q), unit norm, and Eigen's branch choice keeps the result well conditioned
    assert min(np.abs(q2 - q).max(), np.abs(q2 + q).max()) < 1e-12
    assert abs(np.linalg.norm(q2) - 1) < 1e-13


def test_rotation_to_quaternion_at_the_branch_points(L):
    """180-degree rotations about each axis (trace = -1) and the identity: the four branches of Eigen's conversion."""
    for R, want in ((np.diag([1.0, -1, -1]), [1, 0, 0, 0]), (np.diag([-1.0, 1, -1]), [0, 1, 0, 0]),
                    (np.diag([-1.0, -1, 1]), [0, 0, 1, 0]), (np.eye(3), [0, 0, 0, 1])):
        q = np.empty(4)
        L.harness_rot_to_quat(_d(np.ascontiguousarray(R.reshape(-1))), _d(q))
        assert np.abs(np.abs(q) - want).max() < 1e-15


@settings(max_examples=200, deadline=None)
@given(st.lists(finite, min_size=42, max_size=42), st.floats(min_value=1e-8, max_value=10.0))
def test_cholesky_step(L, vals, ridge):
    B = np.array(vals[:36]).reshape(6, 6)
    b = np.array(vals[36:])
    A = B @ B.T + ridge * np.eye(6)
    y = np.empty(6)
    ok = L.harness_chol6_solve(_d(np.ascontiguousarray(A.reshape(-1))), _d(b), _d(y))
    assert ok == 1
    ref = np.linalg.solve(A, b)
    assert np.abs(y - ref).max() <= 1e-9 * np.linalg.cond(A) * max(1.0, np.abs(ref).max()) * 1e-3 + 1e-12
    # not positive definite -> refused (the LM then treats the step as invalid, like Ceres' failed linear solve)
    A[2, 2] = -1.0
    assert L.harness_chol6_solve(_d(np.ascontiguousarray(A.reshape(-1))), _d(b), _d(y)) == 0


@settings(max_examples=200, deadline=None)
@given(st.integers(min_value=1, max_value=8), st.lists(finite, min_size=72, max_size=72))
def test_pivoted_elimination(L, n, vals):
    A = np.array(vals[:n * n]).reshape(n, n) + 2.0 * np.eye(n)
    b = np.array(vals[64:64 + n])
    if np.linalg.cond(A) > 1e8:
        return
    a2, b2 = np.ascontiguousarray(A.reshape(-1)).copy(), b.copy()
    assert L.harness_solve_linear(_d(a2), _d(b2), n) == 1
    np.testing.assert_allclose(b2, np.linalg.solve(A, b), rtol=0, atol=1e-8 * max(1.0, np.abs(b).max()))
    z = np.zeros(n * n)
    assert L.harness_solve_linear(_d(z), _d(b.copy()), n) == 0  # singular


@settings(max_examples=300, deadline=None)
@given(st.floats(min_value=0.0, max_value=1.4), st.tuples(*[st.floats(min_value=-0.03, max_value=0.03)] * 4))
def test_kannala_brandt_inverse(L, theta, ks):
    """theta -> r(theta) -> theta on the monotone branch (the reference solves the degree-9 polynomial by eigenvalues and
    keeps the smallest positive real root, EquidistantCamera.cc:601-672)."""
    k = np.array([363.0, 363.2, 370.1, 240.3, *ks])
    th = np.linspace(0, 1.45, 200)
    r = th * (1 + k[4] * th**2 + k[5] * th**4 + k[6] * th**6 + k[7] * th**8)
    if np.any(np.diff(r) <= 0):
        return  # not a physical lens: r(theta) must grow over the field of view
    rr = L.harness_equi_r(_d(k), theta)
    back = L.harness_equi_theta_from_r(_d(k), rr)
    assert abs(back - theta) < 1e-10

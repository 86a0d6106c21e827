"""LineFittingCeres (reference src/LaseCamCalCeres.cpp:385-433): per-scan robust fit of
m0 x + m1 y + 1 = 0, CauchyLoss(0.05), <= 10 Ceres LM iterations.
CPU: C oracle == numpy twin; the product's 2-parameter LM state machine (host build of csrc/clc_linefit.cuh) driven
by oracle-style sums reproduces the oracle.  GPU: the batched kernel == the oracle scan by scan."""
import ctypes as C

import numpy as np
import pytest


def make_scans(n_scans, seed=0, outliers=True):
    rng = np.random.default_rng(seed)
    scans, truth = [], []
    for _ in range(n_scans):
        m = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1)])
        if np.linalg.norm(m) < 0.2:
            m = m + 0.3
        n = int(rng.integers(3, 400))
        t = rng.uniform(-1, 1, size=n)
        d = np.array([-m[1], m[0]])
        xy = -m / (m @ m) + t[:, None] * d + rng.normal(size=(n, 2)) * 0.01
        if outliers and n > 20:
            idx = rng.choice(n, size=n // 15, replace=False)
            xy[idx] += rng.normal(size=(len(idx), 2)) * 0.3
        scans.append(np.c_[xy, np.zeros(n)])
        truth.append(m)
    return scans, np.array(truth)


def test_oracle_equals_numpy_twin(oracle, oracle_np):
    scans, truth = make_scans(12, seed=1)
    for pts, m in zip(scans, truth):
        for start in ((0.0, 0.0), tuple(1.5 * m), (5.0, -3.0)):
            l, s, tr = oracle.line_fit(pts, start)
            ln, term, trn = oracle_np.line_fit(pts, start)
            assert oracle.TERMINATION[s.termination] == term
            np.testing.assert_allclose(l, ln, rtol=0, atol=1e-12)
            np.testing.assert_allclose([t.cost for t in tr][: len(trn)], [t["cost"] for t in trn], rtol=1e-10)
            if len(pts) > 50 and start == (0.0, 0.0):
                assert np.abs(l - m).max() < 0.02  # robust to the outliers


def line_sums(pts, m, a=0.05):
    x, y = pts[:, 0], pts[:, 1]
    r = m[0] * x + m[1] * y + 1.0
    u = 1.0 + r * r / (a * a)
    w = 1.0 / u
    return np.array([(w * x * x).sum(), (w * x * y).sum(), (w * y * y).sum(), (w * r * x).sum(), (w * r * y).sum(),
                     0.5 * a * a * np.log(u).sum()])


def test_two_parameter_state_machine_reproduces_the_oracle(harness, oracle):
    L = harness.L
    dp = C.POINTER(C.c_double)
    L.harness_lm2_size.restype = C.c_int
    L.harness_lm2_init.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.harness_lm2_update.argtypes = [C.c_void_p, dp, C.c_int]
    L.harness_lm2_done.argtypes = [C.c_void_p]
    L.harness_lm2_get.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    scans, truth = make_scans(20, seed=2)
    for pts, m in zip(scans, truth):
        for start, max_iter in (((0.0, 0.0), 10), (tuple(1.5 * m), 10), ((0.0, 0.0), 2), ((3.0, 7.0), 10)):
            st = C.create_string_buffer(L.harness_lm2_size())
            L.harness_lm2_init(st, start[0], start[1])
            cand, x = np.empty(2), np.empty(2)
            it, sw = C.c_int(), C.c_int()
            while not L.harness_lm2_done(st):
                L.harness_lm2_get(st, cand.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(it), C.byref(sw))
                sums = line_sums(pts, cand.copy())
                L.harness_lm2_update(st, sums.ctypes.data_as(dp), max_iter)
            L.harness_lm2_get(st, cand.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(it), C.byref(sw))
            lo, so, _ = oracle.line_fit(pts, start, max_num_iterations=max_iter)
            assert L.harness_lm2_done(st) == so.termination
            np.testing.assert_allclose(x, lo, rtol=0, atol=1e-10)
            assert sw.value <= max_iter + 1


@pytest.mark.gpu
def test_batched_line_fit_matches_oracle(oracle):
    from camlasercalibratool_b200 import LineFittingCeres, Problem

    scans, truth = make_scans(300, seed=3)
    scans[5] = np.zeros((0, 3))  # an empty scan: the line stays at its start value
    off = np.concatenate([[0], np.cumsum([len(s) for s in scans])])
    pts = np.concatenate(scans)
    fp = np.tile([0, 0, 0, 1, 0, 0, 1.0], (len(scans), 1))
    rng = np.random.default_rng(4)
    starts = np.zeros((len(scans), 2))
    starts[100:200] = truth[100:200] * rng.uniform(0.5, 1.5, size=(100, 1))
    starts[200:] = rng.normal(size=(100, 2)) * 3
    with Problem.from_arrays(fp, off, pts) as g:
        lines, info = g.line_fit(starts)
        lines2, _ = g.line_fit(starts)
        assert np.array_equal(lines, lines2)
        for f, s in enumerate(scans):
            lo, so, tro = oracle.line_fit(s, starts[f])
            np.testing.assert_allclose(lines[f], lo, rtol=0, atol=1e-9, err_msg=f"scan {f}")
            assert int(info[f, 0]) == so.termination, f
            assert abs(info[f, 3] - so.final_cost) <= 1e-9 * max(so.final_cost, 1e-30) + 1e-18
        # iteration limit honoured
        l2, i2 = g.line_fit(starts, max_num_iterations=1)
        lo, so, _ = oracle.line_fit(scans[0], starts[0], max_num_iterations=1)
        np.testing.assert_allclose(l2[0], lo, atol=1e-10)
    # the reference's per-scan call shape
    line = np.zeros(2)
    LineFittingCeres(scans[7], line)
    np.testing.assert_allclose(line, oracle.line_fit(scans[7], (0, 0))[0], atol=1e-9)


def test_two_parameter_state_machine_soak(harness, oracle):
    """300 random scans incl. degenerate ones (1-3 points, all points identical, gross outliers, 1000x scale, odd starts,
    0 / 3 / 50 iteration limits): the device state machine ends for the same reason and at the same line as the oracle.
    (Scans through the sensor origin, where m.p + 1 = 0 has no solution and J is rank deficient, agree only to ~1e-7 --
    normal equations vs QR on a singular problem -- and are left out; a board cannot pass through the laser.)"""
    L = harness.L
    dp = C.POINTER(C.c_double)
    L.harness_lm2_size.restype = C.c_int
    L.harness_lm2_init.argtypes = [C.c_void_p, C.c_double, C.c_double]
    L.harness_lm2_update.argtypes = [C.c_void_p, dp, C.c_int]
    L.harness_lm2_done.argtypes = [C.c_void_p]
    L.harness_lm2_get.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rng = np.random.default_rng(99)
    seen = set()
    for _ in range(300):
        n = int(rng.choice([1, 2, 3, 5, 20, 100, 400]))
        kind = int(rng.integers(0, 4))
        ang, d = rng.uniform(0, 2 * np.pi), rng.uniform(0.3, 5.0)
        t = rng.uniform(-1, 1, size=n) * rng.uniform(0.05, 2.0)
        nrm = np.array([np.cos(ang), np.sin(ang)])
        pts2 = d * nrm + t[:, None] * np.array([-nrm[1], nrm[0]]) + rng.normal(size=(n, 2)) * rng.choice([0.0, 0.002, 0.02])
        if kind == 1:
            pts2[rng.random(n) < 0.2] += rng.normal(size=2)
        elif kind == 2:
            pts2[:] = pts2[0]
        elif kind == 3:
            pts2 *= 1e3
        pts = np.c_[pts2, np.zeros(n)]
        start = [(0.0, 0.0), tuple(rng.normal(size=2) * 3), tuple(-nrm / d)][int(rng.integers(0, 3))]
        max_iter = int(rng.choice([10, 10, 3, 0, 50]))
        st = C.create_string_buffer(L.harness_lm2_size())
        L.harness_lm2_init(st, start[0], start[1])
        cand, x = np.empty(2), np.empty(2)
        it, sw = C.c_int(), C.c_int()
        for _guard in range(200):
            if L.harness_lm2_done(st):
                break
            L.harness_lm2_get(st, cand.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(it), C.byref(sw))
            sums = line_sums(pts, cand.copy())
            L.harness_lm2_update(st, sums.ctypes.data_as(dp), max_iter)
        L.harness_lm2_get(st, cand.ctypes.data_as(dp), x.ctypes.data_as(dp), C.byref(it), C.byref(sw))
        lo, so, _ = oracle.line_fit(pts, start, max_num_iterations=max_iter)
        assert L.harness_lm2_done(st) == so.termination
        np.testing.assert_allclose(x, lo, rtol=0, atol=1e-9 * max(1.0, np.abs(lo).max()))
        seen.add(so.termination)
    assert len(seen) >= 3

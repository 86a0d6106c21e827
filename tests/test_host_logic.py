"""CPU tests of the product's host/device (CLC_HD) logic, compiled with g++ from the very sources the GPU runs
(tests/host_harness.cpp): the on-device LM state machine, the moment -> normal-equation expansion, plane / edge
plane math and the synthetic generator -- each against the oracle.  No CUDA call is made here."""
import ctypes as C

import numpy as np
import pytest

from conftest import pack_sums

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])


def moments_of(points, w):
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    return np.array([w.sum(), (w * x).sum(), (w * y).sum(), (w * z).sum(), (w * x * x).sum(), (w * x * y).sum(),
                     (w * x * z).sum(), (w * y * y).sum(), (w * y * z).sum(), (w * z * z).sum()])


def sweep_by_moments(harness, p, pose7, use_loss=True, a=0.05, split=None):
    """Emulates the kernel on the CPU: per piece weighted moments + log of the cost product, then the product's
    own expand_lm.  `split` cuts every frame into pieces of at most that many points (partial-frame pieces)."""
    out = np.zeros(28)
    dp = harness.dp
    pose7 = np.ascontiguousarray(pose7, dtype=np.float64)
    for f in range(p.n_frames):
        b, e = int(p.offsets[f]), int(p.offsets[f + 1])
        if e <= b:
            continue
        plane = np.empty(4)
        harness.L.harness_frame_plane(dp(np.ascontiguousarray(p.frame_pose[f])), dp(plane))
        m, c = np.empty(3), C.c_double()
        harness.L.harness_frame_consts(dp(plane), dp(pose7), dp(m), C.byref(c))
        step = (e - b) if split is None else split
        for s0 in range(b, e, step):
            pts = p.points[s0:min(e, s0 + step)]
            ee = pts @ m + c.value
            if use_loss:
                u = 1.0 + ee * ee / (a * a)
                w, cost_term = 1.0 / u, float(np.sum(np.log(u)))
            else:
                w, cost_term = np.ones(len(pts)), float(np.sum(ee * ee))
            S = moments_of(pts, w)
            harness.L.harness_expand_lm(dp(plane), dp(pose7), float(e - b), dp(S), int(use_loss), cost_term, a * a, dp(out))
        if p.edge_points is not None:
            p1, p2 = np.empty(4), np.empty(4)
            harness.L.harness_edge_planes(dp(np.ascontiguousarray(p.frame_pose[f])), dp(p1), dp(p2))
            for k, pl in enumerate((p1, p2)):
                pt = p.edge_points[f, 3 * k:3 * k + 3][None, :]
                harness.L.harness_frame_consts(dp(pl), dp(pose7), dp(m), C.byref(c))
                ee = pt @ m + c.value
                u = 1.0 + ee * ee / (a * a) if use_loss else np.ones(1)
                S = moments_of(pt, 1.0 / u)
                ct = float(np.log(u)[0]) if use_loss else float(ee[0] ** 2)
                harness.L.harness_expand_lm(dp(pl), dp(pose7), float(e - b), dp(S), int(use_loss), ct, a * a, dp(out))
    return out


@pytest.mark.parametrize("edges", [False, True])
@pytest.mark.parametrize("use_loss", [True, False])
def test_moment_expansion_equals_direct_accumulation(harness, oracle, edges, use_loss):
    p = oracle.generate(30, 48, seed=4, sigma=0.02, exact_m=True, with_edges=edges, use_loss=use_loss)
    rng = np.random.default_rng(2)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    for pose in (X0, oracle.ground_truth()[1], np.concatenate([rng.normal(size=3) * 0.3, q])):
        cost, H, g = oracle.evaluate_normal(p, pose)
        ref = pack_sums(cost, H, g)
        for split in (None, 7):  # whole frames, and frames cut into partial pieces (linearity of the expansion)
            got = sweep_by_moments(harness, p, pose, use_loss=use_loss, split=split)
            np.testing.assert_allclose(got[:21], ref[:21], rtol=0, atol=2e-13 * np.abs(ref[:21]).max())
            # g = S2 m + c S1 cancels to the size of the residuals: its absolute error scales with |H|, not |g|
            np.testing.assert_allclose(got[21:27], ref[21:27], rtol=0, atol=2e-13 * np.abs(ref[:21]).max())
            assert abs(got[27] - ref[27]) <= 2e-13 * abs(ref[27])


def test_z_nonzero_points(harness, oracle):
    """Oberserve allows z != 0 even though the live callers never produce it."""
    p = oracle.generate(12, 20, seed=9, sigma=0.01, exact_m=True)
    rng = np.random.default_rng(3)
    pts = p.points.copy()
    pts[:, 2] = rng.normal(size=len(pts)) * 0.2
    p2 = oracle.Problem(p.frame_pose, p.offsets, pts)
    pose = oracle.pose_plus(oracle.ground_truth()[1], rng.normal(size=6) * 0.05)
    cost, H, g = oracle.evaluate_normal(p2, pose)
    got = sweep_by_moments(harness, p2, pose)
    np.testing.assert_allclose(got, pack_sums(cost, H, g), rtol=0, atol=2e-13 * np.abs(H).max())


def test_closed_form_expansion(harness, oracle):
    p = oracle.generate(25, 30, seed=6, sigma=0.01, exact_m=True)
    _, _, AtA, Atb = oracle.closed_form(p)
    out = np.zeros(54)
    for f in range(p.n_frames):
        plane = np.empty(4)
        harness.L.harness_frame_plane(harness.dp(np.ascontiguousarray(p.frame_pose[f])), harness.dp(plane))
        pts = p.points[p.offsets[f]:p.offsets[f + 1]]
        harness.L.harness_expand_closed(harness.dp(plane), harness.dp(moments_of(pts, np.ones(len(pts)))), harness.dp(out))
    iu = np.triu_indices(9)
    np.testing.assert_allclose(out[:45], AtA[iu], rtol=0, atol=1e-12 * np.abs(AtA).max())
    np.testing.assert_allclose(out[45:], Atb, rtol=0, atol=1e-12 * np.abs(Atb).max())


@pytest.mark.parametrize("case", ["noise_free", "noisy", "edges", "no_loss", "bad_start"])
def test_device_lm_state_machine_reproduces_the_ceres_trajectory(harness, oracle, case):
    """lm_update (the code thread 0 runs on the GPU) driven by oracle sums == oracle_solve (Ceres restatement with
    Householder QR on the materialised Jacobian): same termination, same accept/reject sequence, same costs."""
    kw = dict(noise_free=dict(sigma=0.0), noisy=dict(sigma=0.01), edges=dict(sigma=0.01, with_edges=True, exact_m=True),
              no_loss=dict(sigma=0.01, use_loss=False), bad_start=dict(sigma=0.02))[case]
    p = oracle.generate(50, 180, seed=2, **kw)
    x0 = X0
    if case == "bad_start":  # a start that forces rejected steps
        x0 = np.array([3.0, -2.0, 4.0, 0.7, 0.1, -0.7, 0.1])
        x0[3:] /= np.linalg.norm(x0[3:])

    def sums(pose):
        c, H, g = oracle.evaluate_normal(p, pose)
        return pack_sums(c, H, g)

    x, done, trace, sweeps = harness.lm_run(sums, x0)
    xo, so, tro = oracle.solve(p, x0)  # DENSE_QR oracle
    assert done == so.termination
    assert len(trace) == so.num_iterations
    ang, dt = oracle.pose_error(x, xo)
    assert ang < 1e-9 and dt < 1e-9
    for a, b in zip(trace, tro):
        assert (a.iteration, a.step_is_valid, a.step_is_successful) == (b.iteration, b.step_is_valid, b.step_is_successful)
        assert abs(a.cost - b.cost) <= 1e-9 * max(abs(b.cost), 1e-30) + 1e-18
        assert abs(a.trust_region_radius - b.trust_region_radius) <= 1e-6 * b.trust_region_radius
    # one sweep per LM iteration (speculative evaluation): never more sweeps than iterations
    assert sweeps <= len(trace)
    if case == "bad_start":
        assert any(not t.step_is_successful for t in trace)


def test_lm_iteration_limit_and_invalid_inputs(harness, oracle):
    p = oracle.generate(50, 180, seed=1, sigma=0.01)

    def sums(pose):
        c, H, g = oracle.evaluate_normal(p, pose)
        return pack_sums(c, H, g)

    x, done, trace, _ = harness.lm_run(sums, X0, harness.default_options(max_num_iterations=3))
    xo, so, _ = oracle.solve(p, X0, oracle.default_options(max_num_iterations=3))
    assert done == so.termination == 5 and len(trace) == so.num_iterations == 4
    assert oracle.pose_error(x, xo)[0] < 1e-10
    # max_num_iterations = 0: only iteration 0 is evaluated
    x, done, trace, sweeps = harness.lm_run(sums, X0, harness.default_options(max_num_iterations=0))
    assert done == 5 and len(trace) == 1 and sweeps == 1 and np.array_equal(x, X0)
    # NaN at the start point (garbage input): FAILURE, as Ceres' evaluator reports
    x, done, trace, _ = harness.lm_run(lambda pose: np.full(28, np.nan), X0)
    assert done == 6 and np.array_equal(x, X0)
    # NaN at candidates only: every step is "a step with infinite cost" -> rejected until the radius collapses
    first = [True]

    def sums_nan_later(pose):
        if first[0]:
            first[0] = False
            return sums(pose)
        return np.full(28, np.nan)

    x, done, trace, _ = harness.lm_run(sums_nan_later, X0)
    assert np.array_equal(x, X0) and done in (2, 4, 5) and not any(t.step_is_successful for t in trace[1:])


def test_plane_and_plus_match_oracle(harness, oracle):
    rng = np.random.default_rng(8)
    for _ in range(50):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        fp = np.concatenate([q * rng.uniform(0.8, 1.2), rng.normal(size=3)])
        pl, p1, p2 = np.empty(4), np.empty(4), np.empty(4)
        harness.L.harness_frame_plane(harness.dp(fp), harness.dp(pl))
        harness.L.harness_edge_planes(harness.dp(fp), harness.dp(p1), harness.dp(p2))
        np.testing.assert_allclose(pl, oracle.frame_plane(fp), rtol=1e-13, atol=1e-15)
        a, b = oracle.edge_planes(fp)
        np.testing.assert_allclose(p1, a, rtol=1e-13, atol=1e-15)
        np.testing.assert_allclose(p2, b, rtol=1e-13, atol=1e-15)
        x = np.concatenate([rng.normal(size=3), q])
        d = rng.normal(size=6) * 0.3
        xp = np.empty(7)
        harness.L.harness_pose_plus(harness.dp(x), harness.dp(d), harness.dp(xp))
        np.testing.assert_allclose(xp, oracle.pose_plus(x, d), rtol=1e-14, atol=1e-16)


def test_generator_twin(harness, oracle):
    """The device generator's code (exact-M mode) == the oracle's generator: identical Philox stream, identical
    accept/redraw decisions, points equal to rounding."""
    out = (C.c_uint32 * 4)()
    for seed, lo, hi in ((0, 0, 0), (1, 2, 3), (2**63 + 5, 2**40 + 7, (1 << 56) | 9)):
        harness.L.harness_philox(seed, lo, hi, out)
        assert [int(v) for v in out] == oracle.philox(seed, lo, hi)
    # Philox4x32-10 known answer (Random123 kat_vectors: counter = key = 0)
    harness.L.harness_philox(0, 0, 0, out)
    assert [hex(v) for v in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    for edges in (False, True):
        p = oracle.generate(64, 24, seed=13, sigma=0.01, exact_m=True, with_edges=edges)
        for f in range(64):
            fp = np.empty(7)
            harness.L.harness_gen_frame_pose(13, f, int(edges), harness.dp(fp))
            np.testing.assert_allclose(fp, p.frame_pose[f], rtol=0, atol=1e-15)
            pts = np.empty((24, 3))
            harness.L.harness_gen_points(13, 0.01, f, 24, harness.dp(fp), harness.dp(pts))
            np.testing.assert_allclose(pts, p.points[p.offsets[f]:p.offsets[f + 1]], rtol=0, atol=1e-13)
            if edges:
                ep = np.empty(6)
                assert harness.L.harness_gen_edge_points(harness.dp(fp), harness.dp(ep)) == 1
                np.testing.assert_allclose(ep, p.edge_points[f], rtol=0, atol=1e-13)


def test_state_machine_soak_on_random_problems(harness, oracle):
    """80 random problems (sizes, noise, loss on/off, edge residuals, starts from exact to far off): the device LM code
    (Cholesky on the normal equations) takes the same accept/reject decisions and stops for the same reason as the
    Ceres-shaped oracle (Householder QR on the materialised Jacobian).  A 400-problem run of the same loop was clean."""
    rng = np.random.default_rng(12345)
    seen = set()
    for _ in range(80):
        n_frames, beams = int(rng.integers(5, 60)), int(rng.integers(20, 200))
        sigma = float(rng.choice([0.0, 0.005, 0.01, 0.03]))
        edges, loss = bool(rng.random() < 0.3), bool(rng.random() < 0.8)
        p = oracle.generate(n_frames, beams, seed=int(rng.integers(1, 10**6)), sigma=sigma, with_edges=edges, exact_m=edges,
                            use_loss=loss)
        if p.n_points == 0:
            continue
        scale = float(rng.choice([0.0, 0.05, 0.5, 2.0]))
        x0 = X0
        if scale > 0:
            if rng.random() < 0.7:
                x0 = oracle.pose_plus(oracle.ground_truth()[1], rng.normal(size=6) * scale)
            else:
                q = rng.normal(size=4)
                x0 = np.concatenate([rng.normal(size=3) * scale, q / np.linalg.norm(q)])

        def sums(pose):
            c, H, g = oracle.evaluate_normal(p, pose)
            return pack_sums(c, H, g)

        x, done, trace, sweeps = harness.lm_run(sums, x0)
        xo, so, tro = oracle.solve(p, x0)
        assert done == so.termination and len(trace) == so.num_iterations
        assert all((a.step_is_valid, a.step_is_successful) == (b.step_is_valid, b.step_is_successful) for a, b in zip(trace, tro))
        ang, dt = oracle.pose_error(x, xo)
        assert ang < 1e-6 and dt < 1e-6
        seen.add(so.termination)
    assert len(seen) >= 3  # function / parameter / gradient tolerance (and sometimes the iteration limit) all occur


@pytest.mark.parametrize("edges,use_loss", [(False, True), (True, True), (True, False)])
def test_direct_residual_accumulation_of_the_one_cluster_kernel(harness, oracle, edges, use_loss):
    """csrc/clc_small.cuh adds every residual directly to the 28 sums (accumulate_residual, compiled here for the host from
    the same source): summed over all residuals -- points with their frame's plane and 1/#points, edge residuals with their
    edge plane -- it must give the oracle's (cost, H, g)."""
    p = oracle.generate(30, 48, seed=4, sigma=0.02, exact_m=True, with_edges=edges, use_loss=use_loss)
    rng = np.random.default_rng(5)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    a2 = p.cauchy_a ** 2
    for pose in (X0, oracle.ground_truth()[1], np.concatenate([rng.normal(size=3) * 0.3, q])):
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        acc = np.zeros(28)
        for f in range(p.n_frames):
            plane = np.ascontiguousarray(oracle.frame_plane(p.frame_pose[f]), dtype=np.float64)
            b, e = int(p.offsets[f]), int(p.offsets[f + 1])
            for i in range(b, e):
                xyz = np.ascontiguousarray(p.points[i], dtype=np.float64)
                harness.L.harness_accumulate_residual(harness.dp(plane), harness.dp(pose), harness.dp(xyz), float(e - b),
                                                      int(use_loss), a2, harness.dp(acc))
            if edges and e > b:
                p1, p2 = oracle.edge_planes(p.frame_pose[f])
                for plane_e, xyz in ((p1, p.edge_points[f, :3]), (p2, p.edge_points[f, 3:])):
                    plane_e = np.ascontiguousarray(plane_e, dtype=np.float64)
                    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
                    harness.L.harness_accumulate_residual(harness.dp(plane_e), harness.dp(pose), harness.dp(xyz), float(e - b),
                                                          int(use_loss), a2, harness.dp(acc))
        cost, H, g = oracle.evaluate_normal(p, pose)
        ref = pack_sums(cost, H, g)
        np.testing.assert_allclose(acc[:27], ref[:27], rtol=0, atol=2e-13 * np.abs(ref[:21]).max())
        assert abs(acc[27] - ref[27]) <= 2e-13 * abs(ref[27])

"""GPU tests (-m gpu) of the ways the LM loop is driven (csrc/clc_api.cu solve_all, CLC_LOOP_IN_KERNEL / CLC_SMALL_KERNEL):
  0  one launch per LM iteration, chained with programmatic dependent launch
  1  (default) small problems run the whole LM loop inside ONE launch -- up to 16384 residuals in the one-cluster kernel of
     csrc/clc_small.cuh (residuals resident in registers); with CLC_SMALL_KERNEL=0 in the LOOP instantiation of the sweep kernel
  2  every problem does: a persistent grid, block 0 hands the next pose to the other blocks as tagged words
The sweep-kernel drivers share the data path (partition, per-block sums, gather order), so their trajectories must agree bit for
bit; the one-cluster kernel adds the same residuals in another order (agreement to 1e-12, identical decisions); and each must
agree with the oracle.  Also: the per-scan LineFittingCeres path (one warp on the scan's own AoS array) equals the
batched kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])


def _solve(monkeypatch, mode, make):
    monkeypatch.setenv("CLC_LOOP_IN_KERNEL", str(mode))
    with make() as g:
        x, s, tr = g.solve(X0)
        x2, s2, tr2 = g.solve(X0)  # a second solve on the same problem: sequence numbers and mailboxes carry over
        assert np.array_equal(x, x2) and [t.cost for t in tr] == [t.cost for t in tr2]
        near = g.solve(x)  # a start at the solution: terminates after very few sweeps
        cost, H, grad = g.eval(x)  # a plain single-sweep launch after a looping one
    return x, s, [t.cost for t in tr], [t.trust_region_radius for t in tr], near[1].num_iterations, cost


@pytest.mark.parametrize("shape", ["config1", "one_block_edges", "multi_block", "multi_block_planar", "no_loss"])
def test_loop_modes_agree_bit_for_bit(oracle, monkeypatch, shape):
    from camlasercalibratool_b200 import Problem

    monkeypatch.setenv("CLC_PLANAR_MIN_POINTS", "0")
    monkeypatch.setenv("CLC_PLANAR", "1" if shape == "multi_block_planar" else "0")
    if shape == "config1":
        p = oracle.generate(50, 180, seed=1, sigma=0.01)
    elif shape == "one_block_edges":
        p = oracle.generate(30, 400, seed=4, sigma=0.01, exact_m=True, with_edges=True)
    elif shape == "no_loss":
        p = oracle.generate(40, 250, seed=6, sigma=0.01, exact_m=True, use_loss=False)
    else:
        p = oracle.generate(300, 700, seed=5, sigma=0.01, exact_m=True)  # 210k points: ~100 blocks of 16 warps

    def make():
        return Problem.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points, use_loss=p.use_loss, cauchy_a=p.cauchy_a)

    monkeypatch.setenv("CLC_SMALL_KERNEL", "0")
    ref = _solve(monkeypatch, 0, make)
    for mode in (1, 2):  # the sweep kernel's drivers: bit for bit
        got = _solve(monkeypatch, mode, make)
        assert np.array_equal(got[0], ref[0]), (shape, mode)
        assert got[2] == ref[2] and got[3] == ref[3], (shape, mode)
        assert got[1].termination == ref[1].termination and got[1].num_iterations == ref[1].num_iterations
        assert got[1].num_sweeps == ref[1].num_sweeps and got[4] == ref[4] and got[5] == ref[5]
    if p.n_points + (0 if p.edge_points is None else 2 * p.n_frames) <= 16384:
        monkeypatch.setenv("CLC_SMALL_KERNEL", "1")  # the one-cluster kernel: same decisions, sums in another order
        got = _solve(monkeypatch, 1, make)
        np.testing.assert_allclose(got[0], ref[0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(got[2], ref[2], rtol=1e-12)
        assert got[1].termination == ref[1].termination and got[1].num_iterations == ref[1].num_iterations
        assert got[1].num_sweeps == ref[1].num_sweeps and got[4] == ref[4]
    xo, so, _ = oracle.solve(p, X0)
    ang, dt = oracle.pose_error(ref[0], xo)
    assert ang < 1e-6 and dt < 1e-6 and ref[1].termination == so.termination and ref[1].num_iterations == so.num_iterations


def test_looping_launch_respects_max_num_iterations(oracle, monkeypatch):
    from camlasercalibratool_b200 import Problem, default_options

    p = oracle.generate(50, 180, seed=2, sigma=0.01)
    for mode in (0, 1, 2):
        monkeypatch.setenv("CLC_LOOP_IN_KERNEL", str(mode))
        with Problem.from_arrays(p.frame_pose, p.offsets, p.points) as g:
            for cap in (0, 1, 3):
                x, s, tr = g.solve(X0, default_options(max_num_iterations=cap))
                xo, so, _ = oracle.solve(p, X0, oracle.default_options(max_num_iterations=cap))
                assert s.termination == so.termination and s.num_iterations == so.num_iterations, (mode, cap)
                ang, dt = oracle.pose_error(x, xo)
                assert ang < 1e-9 and dt < 1e-9


def test_per_scan_line_fit_equals_the_batched_kernel(oracle):
    from camlasercalibratool_b200 import LineFittingCeres, Problem

    rng = np.random.default_rng(8)
    p = oracle.generate(40, 300, seed=12, sigma=0.01)
    pts = p.points.copy()
    idx = rng.choice(len(pts), size=len(pts) // 30, replace=False)
    pts[idx, :2] += rng.normal(size=(len(idx), 2)) * 0.3  # outliers: the reason the fit is robust
    with Problem.from_arrays(p.frame_pose, p.offsets, pts) as g:
        lines, info = g.line_fit()
    for f in range(0, p.n_frames, 3):
        scan = pts[p.offsets[f]:p.offsets[f + 1]]
        line = np.zeros(2)
        LineFittingCeres(scan, line)
        assert np.array_equal(line, lines[f]), f  # the same arithmetic on the same points: identical bits
        ref = oracle.line_fit(scan)[0]
        np.testing.assert_allclose(line, ref, rtol=0, atol=1e-9)
    line = np.array([0.3, -0.2])
    LineFittingCeres(np.zeros((0, 3)), line)  # an empty scan leaves a finite line (the reference would not even get here)
    assert np.all(np.isfinite(line))


def test_l2_persistence_window_changes_no_bit(oracle, monkeypatch):
    """CLC_L2_PERSIST_MB keeps a share of the coordinate arrays resident in L2 across the LM iterations of a solve: a cache
    policy, not arithmetic -- the trajectory must be identical, and evaluations outside a solve are untouched."""
    from camlasercalibratool_b200 import Problem

    monkeypatch.setenv("CLC_PLANAR_MIN_POINTS", "0")
    out = []
    for mb in ("0", "32"):
        monkeypatch.setenv("CLC_L2_PERSIST_MB", mb)
        with Problem.synthetic(400, 700, seed=5, sigma=0.01) as g:
            x, s, tr = g.solve(X0)
            x2, s2, tr2 = g.solve(X0)
            assert np.array_equal(x, x2)
            out.append((x, [t.cost for t in tr], g.eval(x)[0], s.num_sweeps))
    assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1] and out[0][2] == out[1][2] and out[0][3] == out[1][3]


@pytest.mark.parametrize("case", ["ragged_with_empty_frames", "edges", "no_loss", "planar_forced", "off_plane", "full_16384", "noise_free"])
def test_one_cluster_kernel_follows_the_oracle(oracle, monkeypatch, case):
    """csrc/clc_small.cuh against the oracle on its own: shapes the reference produces (ragged frames, empty frames from the
    validity filter, edge residuals, z != 0) and the size limit."""
    from camlasercalibratool_b200 import Problem, launch_count

    monkeypatch.setenv("CLC_LOOP_IN_KERNEL", "1")
    monkeypatch.setenv("CLC_SMALL_KERNEL", "1")
    monkeypatch.setenv("CLC_PLANAR_MIN_POINTS", "0" if case == "planar_forced" else "1000000000")
    rng = np.random.default_rng(3)
    if case == "ragged_with_empty_frames":
        base = oracle.generate(60, 180, seed=2, sigma=0.01)  # faithful mode: ragged
        counts = np.diff(base.offsets)
        keep = rng.integers(0, 4, size=60) != 0  # drop a quarter of the frames' points entirely
        frames = [base.points[base.offsets[f]:base.offsets[f + 1]] if keep[f] else np.zeros((0, 3)) for f in range(60)]
        off = np.concatenate([[0], np.cumsum([len(f) for f in frames])])
        p = oracle.Problem(base.frame_pose, off, np.concatenate(frames, axis=0))
    elif case == "edges":
        p = oracle.generate(40, 150, seed=4, sigma=0.01, exact_m=True, with_edges=True)
    elif case == "no_loss":
        p = oracle.generate(50, 180, seed=5, sigma=0.01, use_loss=False)
    elif case == "planar_forced":
        p = oracle.generate(50, 180, seed=6, sigma=0.01)
    elif case == "off_plane":
        b = oracle.generate(50, 180, seed=7, sigma=0.01)
        p = oracle.Problem(b.frame_pose, b.offsets, b.points + np.array([0, 0, 1.0]) * rng.normal(scale=0.2, size=(b.n_points, 1)))
    elif case == "full_16384":
        p = oracle.generate(64, 256, seed=8, sigma=0.01, exact_m=True)  # exactly the limit
    else:
        b = oracle.generate(30, 100, seed=9, sigma=0.0, exact_m=True)
        p = oracle.Problem(b.frame_pose, b.offsets, b.points)
    with Problem.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points, use_loss=p.use_loss, cauchy_a=p.cauchy_a) as g:
        for x0 in (X0, oracle.pose_plus(oracle.ground_truth()[1], np.array([0.05, -0.04, 0.03, 0.02, -0.03, 0.025]))):
            n0 = launch_count()
            x, s, tr = g.solve(x0)
            assert launch_count() - n0 == 1, "one launch per solve"
            xo, so, tro = oracle.solve(p, x0)
            ang, dt = oracle.pose_error(x, xo)
            assert ang < 1e-6 and dt < 1e-6, (case, ang, dt)
            assert s.termination == so.termination and s.num_iterations == so.num_iterations, (case, s.termination, so.termination)
            # (absolute floor: on noise-free data the cost falls to ~1e-20, where the summation order decides the digits)
            np.testing.assert_allclose([t.cost for t in tr], [t.cost for t in tro][: len(tr)], rtol=1e-9, atol=1e-14 * tro[0].cost)
            x2, s2, _ = g.solve(x0)
            assert np.array_equal(x, x2)  # bit-reproducible


def test_problems_above_the_limit_take_the_streaming_path(oracle, monkeypatch):
    from camlasercalibratool_b200 import Problem, launch_count

    monkeypatch.setenv("CLC_LOOP_IN_KERNEL", "1")
    p = oracle.generate(65, 256, seed=8, sigma=0.01, exact_m=True)  # 16640 residuals: one more frame than the limit
    with Problem.from_arrays(p.frame_pose, p.offsets, p.points) as g:
        n0 = launch_count()
        x, s, _ = g.solve(X0)
        assert launch_count() - n0 > 1
        xo, so, _ = oracle.solve(p, X0)
        ang, dt = oracle.pose_error(x, xo)
        assert ang < 1e-6 and dt < 1e-6 and s.num_iterations == so.num_iterations

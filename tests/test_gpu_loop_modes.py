"""GPU tests (-m gpu) of the three ways the LM loop is driven (csrc/clc_api.cu solve_all, CLC_LOOP_IN_KERNEL):
  0  one launch per LM iteration, chained with programmatic dependent launch
  1  (default) problems that fit one block run the whole LM loop inside ONE launch
  2  every problem does: a persistent grid, block 0 hands the next pose to the other blocks as tagged words
The data path (partition, per-block sums, gather order) is the same in all three, so the trajectories must agree bit for bit,
and each must agree with the oracle.  Also: the per-scan LineFittingCeres path (one warp on the scan's own AoS array) equals the
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

    ref = _solve(monkeypatch, 0, make)
    for mode in (1, 2):
        got = _solve(monkeypatch, mode, make)
        assert np.array_equal(got[0], ref[0]), (shape, mode)
        assert got[2] == ref[2] and got[3] == ref[3], (shape, mode)
        assert got[1].termination == ref[1].termination and got[1].num_iterations == ref[1].num_iterations
        assert got[1].num_sweeps == ref[1].num_sweeps and got[4] == ref[4] and got[5] == ref[5]
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

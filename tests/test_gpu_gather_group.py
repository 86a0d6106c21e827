"""GPU tests (-m gpu) of the upload pipeline and of the in-process multi-GPU group (include/clc_b200.h: clc_problem_create_gather,
clc_group_*): what arrives in HBM equals what was handed over -- per-frame pageable arrays, ragged/empty frames, planar and
non-planar data, a z != 0 that shows up only late -- and a group of G devices solves like one device and like the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _frames_of(p):
    return [p.points[p.offsets[f]:p.offsets[f + 1]].copy() for f in range(p.n_frames)]


@pytest.mark.parametrize("chunk_points", [1024, 5000, 1 << 18])
@pytest.mark.parametrize("nonplanar_at", [None, 0, "last", "middle"])
def test_gather_roundtrip(oracle, monkeypatch, chunk_points, nonplanar_at):
    """Several chunks, several pack threads, chunk boundaries inside frames; the z stream appears only when it must."""
    from camlasercalibratool_b200 import Problem, upload_stats

    monkeypatch.setenv("CLC_UPLOAD_CHUNK_POINTS", str(chunk_points))
    rng = np.random.default_rng(5)
    base = oracle.generate(300, 97, seed=6, sigma=0.01, exact_m=True)
    counts = rng.choice([0, 1, 2, 40, 97, 97, 97], size=300)
    frames = [base.points[base.offsets[f]:base.offsets[f] + c].copy() for f, c in enumerate(counts)]
    P = int(counts.sum())
    if nonplanar_at is not None:
        idx = {0: 0, "last": P - 1, "middle": P // 2}[nonplanar_at]
        f = int(np.searchsorted(np.cumsum(counts), idx, side="right"))
        frames[f][idx - int(np.cumsum(counts)[f] - counts[f]), 2] = 0.125
    with Problem.from_frames(base.frame_pose, frames) as g:
        st = upload_stats()
        d = g.download()
        np.testing.assert_array_equal(d["points"], np.concatenate(frames, axis=0))
        np.testing.assert_array_equal(d["offsets"], np.concatenate([[0], np.cumsum(counts)]))
        np.testing.assert_array_equal(d["frame_pose"], base.frame_pose)
        assert st["chunks"] == -(-P // chunk_points) and not st["direct"] and st["pack_threads"] >= 1
        if nonplanar_at is None:
            assert st["bytes_h2d"] == 16 * P  # planar data travels as x,y only
        else:
            assert 16 * P < st["bytes_h2d"] <= 24 * P
        p = oracle.Problem(base.frame_pose, d["offsets"], d["points"])
        cost, H, grad = g.eval(X0)
        rc, rH, rg = oracle.evaluate_normal(p, X0)
        assert abs(cost - rc) <= 1e-11 * rc and np.abs(H - rH).max() <= 1e-11 * np.abs(rH).max()


def test_flat_pageable_and_pinned_sources_agree(oracle):
    from camlasercalibratool_b200 import Problem, upload_stats
    from camlasercalibratool_b200.api import pinned_array

    p = oracle.generate(200, 300, seed=2, sigma=0.01, exact_m=True)
    with Problem.from_arrays(p.frame_pose, p.offsets, p.points) as a:
        sa = upload_stats()
        ca = a.eval(X0)
        da = a.download()
    pin = pinned_array(p.points.shape)
    pin.array[...] = p.points
    with Problem.from_arrays(p.frame_pose, p.offsets, pin.array) as b:
        sb = upload_stats()
        cb = b.eval(X0)
        db = b.download()
    pin.free()
    assert not sa["direct"] and sb["direct"]
    assert sa["bytes_h2d"] == 16 * p.n_points and sb["bytes_h2d"] == 24 * p.n_points
    np.testing.assert_array_equal(da["points"], db["points"])
    assert ca[0] == cb[0] and np.array_equal(ca[1], cb[1])


def test_group_of_one_is_a_problem(oracle):
    from camlasercalibratool_b200 import Group, Problem

    p = oracle.generate(50, 180, seed=1, sigma=0.01)
    with Group.from_arrays(p.frame_pose, p.offsets, p.points) as g, Problem.from_arrays(p.frame_pose, p.offsets, p.points) as q:
        assert g.sizes() == (1, 50, p.n_points)
        assert g.eval(X0)[0] == q.eval(X0)[0]
        xg, sg, _ = g.solve(X0)
        xq, sq, _ = q.solve(X0)
        assert np.array_equal(xg, xq) and sg.num_iterations == sq.num_iterations
        np.testing.assert_array_equal(g.information(xg)[0], q.information(xq)[0])
        np.testing.assert_array_equal(g.closed_form()[0], q.closed_form()[0])
    xo, so, _ = oracle.solve(p, X0)
    ang, dt = oracle.pose_error(xg, xo)
    assert ang < 1e-6 and dt < 1e-6 and sg.termination == so.termination


@pytest.mark.parametrize("G", [2, 4, 8])
def test_group_matches_single_device_and_oracle(oracle, G):
    """One process, G devices, ragged frames sharded by point count, edges on: the collective eval / solve / information /
    closed form equal the single-device results (summation order only) and the oracle's solve."""
    if _n_gpus() < G:
        pytest.skip(f"needs {G} GPUs")
    from camlasercalibratool_b200 import Group, Problem

    p = oracle.generate(600, 180, seed=4, sigma=0.01, with_edges=True)  # faithful generator: ragged frames
    devices = list(range(G))
    with Group.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points, devices=devices) as g, \
            Problem.from_arrays(p.frame_pose, p.offsets, p.points, p.edge_points, device=0) as q:
        n, nf, npts = g.sizes()
        assert (n, nf, npts) == (G, 600, p.n_points)
        shard_points = [g.problem(i).sizes()[1] for i in range(G)]
        assert sum(shard_points) == p.n_points and max(shard_points) - min(shard_points) <= 2 * 180
        cg, Hg, gg = g.eval(X0)
        cq, Hq, gq = q.eval(X0)
        scale = np.abs(Hq).max()
        assert abs(cg - cq) <= 1e-12 * cq
        np.testing.assert_allclose(Hg, Hq, rtol=0, atol=1e-12 * scale)
        np.testing.assert_allclose(gg, gq, rtol=0, atol=1e-12 * scale)
        xg, sg, trg = g.solve(X0)
        xq, sq, trq = q.solve(X0)
        assert sg.termination == sq.termination and sg.num_iterations == sq.num_iterations
        np.testing.assert_allclose([t.cost for t in trg], [t.cost for t in trq], rtol=1e-10)
        xg2, _, _ = g.solve(X0)  # mailboxes and sequence numbers are reused
        assert np.array_equal(xg, xg2)
        np.testing.assert_allclose(g.information(xg)[0], q.information(xq)[0], rtol=0, atol=1e-11 * scale)
        np.testing.assert_allclose(g.closed_form()[0], q.closed_form()[0], atol=1e-9)
    xo, so, _ = oracle.solve(p, X0)
    ang, dt = oracle.pose_error(xg, xo)
    assert ang < 1e-6 and dt < 1e-6 and sg.termination == so.termination and sg.num_iterations == so.num_iterations


@pytest.mark.parametrize("G", [2, 8])
def test_group_synthetic_recovers_ground_truth(oracle, G):
    if _n_gpus() < G:
        pytest.skip(f"needs {G} GPUs")
    from camlasercalibratool_b200 import Group

    with Group.synthetic(8000, 700, seed=9, sigma=0.0, devices=list(range(G))) as g:
        x, s, _ = g.solve(X0)
    assert s.termination in (1, 2, 3)
    np.testing.assert_allclose(np.linalg.inv(oracle.pose7_to_T(x)), oracle.ground_truth()[0], atol=1e-9)


def test_dropin_rejects_duplicate_devices():
    from camlasercalibratool_b200 import ClcError, Group

    with pytest.raises(ClcError):
        Group.synthetic(100, 100, devices=[0, 0])

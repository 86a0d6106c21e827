"""The N > 1 path on CPU (gloo, world_size 2): frames sharded by rank with clc_shard_range, every rank sweeps its own
shard, the 28 sums are all-reduced, and every rank runs the SAME LM state machine (the product's lm_update compiled
for the host) redundantly -- exactly the structure of the multi-GPU solve, with the oracle standing in for the sweep
kernel and gloo for NCCL.  Checks: ranks stay in lock-step (identical iterates), and the result equals the
single-rank solve.  The GPU version of the same test (NCCL, real kernels) is test_multi_gpu_* in -m gpu."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, harness_path, out_dir, ragged):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from camlasercalibratool_b200 import shard_range
    from conftest import Harness, pack_sums
    from oracle import oracle as O

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = O.generate(60, 90, seed=5, sigma=0.01, exact_m=not ragged)
    b, e = shard_range(full.n_frames, world, rank, full.offsets if ragged else None)
    off = full.offsets[b:e + 1] - full.offsets[b]
    shard = O.Problem(full.frame_pose[b:e], off, full.points[full.offsets[b]:full.offsets[e]])
    h = Harness(harness_path)
    iterates = []

    def sums(pose):
        iterates.append(pose.copy())
        c, H, g = O.evaluate_normal(shard, pose)
        t = torch.from_numpy(pack_sums(c, H, g).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the 28-double all-reduce of the multi-GPU solve
        return t.numpy()

    x, done, trace, sweeps = h.lm_run(sums, np.array([0, 0, 0, 0, 0, 0, 1.0]))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x=x, done=done, iterates=np.array(iterates),
             costs=np.array([t.cost for t in trace]), shard=np.array([b, e]))
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_two_rank_solve_matches_single_rank(harness, oracle, tmp_path, ragged):
    import torch.multiprocessing as mp

    # the session fixture's library path is needed by the workers
    harness_path = harness.L._name
    port = _free_port()
    mp.spawn(_worker, args=(2, port, harness_path, str(tmp_path), ragged), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    # shards tile the frames
    assert r0["shard"][0] == 0 and r0["shard"][1] == r1["shard"][0] and r1["shard"][1] == 60
    # lock-step: both ranks evaluated exactly the same sequence of poses and took the same decisions
    assert np.array_equal(r0["iterates"], r1["iterates"])
    assert np.array_equal(r0["x"], r1["x"]) and r0["done"] == r1["done"]
    assert np.array_equal(r0["costs"], r1["costs"])
    # and the sharded solve is the single-rank solve (sums differ only by association order)
    full = oracle.generate(60, 90, seed=5, sigma=0.01, exact_m=not ragged)
    xo, so, tro = oracle.solve(full, np.array([0, 0, 0, 0, 0, 0, 1.0]))
    ang, dt = oracle.pose_error(r0["x"], xo)
    assert ang < 1e-9 and dt < 1e-9 and int(r0["done"]) == so.termination
    np.testing.assert_allclose(r0["costs"], [t.cost for t in tro], rtol=1e-9)

"""Multi-GPU parity (-m gpu, needs >= 2 GPUs; skipped otherwise): one process per GPU, frames sharded by rank, the 28
normal-equation sums all-reduced with NCCL after every sweep, the LM update run redundantly on every device.
The sharded solve must equal the single-GPU solve and the CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from camlasercalibratool_b200 import Comm, Problem, comm_unique_id, shard_range

    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = Comm(uid[0], world, rank, device=rank)

    def all_gather(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    comm.enable_p2p(all_gather)  # fused in-kernel exchange over NVLink peer memory
    n_frames, beams = 4000, 500
    b, e = shard_range(n_frames, world, rank)
    x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
    res = {}
    with Problem.synthetic(n_frames, beams, seed=3, sigma=0.01, frame_begin=b, frame_end=e, device=rank) as p:
        p.attach_comm(comm)
        for mode, tag in ((1, ""), (0, "_nccl")):  # 1: fused peer exchange (default), 0: ncclAllReduce between kernels
            p.set_allreduce_mode(mode)
            cost, H, g = p.eval(x0)  # all-reduced over the ranks
            x, s, tr = p.solve(x0)
            x2, s2, _ = p.solve(x0)  # a second solve re-uses mailboxes and sequence numbers
            assert np.array_equal(x, x2) and s.num_iterations == s2.num_iterations
            Hi, bi, chi, sv = p.information(x)
            T, un, AtA, Atb = p.closed_form()
            res.update({"cost" + tag: cost, "H" + tag: H, "g" + tag: g, "x" + tag: x, "term" + tag: s.termination,
                        "iters" + tag: s.num_iterations, "costs" + tag: np.array([t.cost for t in tr]), "Hi" + tag: Hi,
                        "chi" + tag: chi, "T" + tag: T, "sweeps" + tag: s.num_sweeps})
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_process_solve_matches_single_gpu(oracle, tmp_path, world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    from camlasercalibratool_b200 import Problem

    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    for k in range(1, world):  # every rank holds the identical all-reduced result and took the same decisions
        for key in ("cost", "H", "g", "x", "costs", "Hi", "T", "cost_nccl", "H_nccl", "x_nccl", "costs_nccl"):
            assert np.array_equal(r[0][key], r[k][key]), key
    # the NCCL path and the fused peer path agree (different association order only)
    ang, dt = oracle.pose_error(r[0]["x"], r[0]["x_nccl"])
    assert ang < 1e-10 and dt < 1e-10 and int(r[0]["iters"]) == int(r[0]["iters_nccl"])
    np.testing.assert_allclose(r[0]["H"], r[0]["H_nccl"], rtol=0, atol=1e-12 * np.abs(r[0]["H"]).max())
    x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
    with Problem.synthetic(4000, 500, seed=3, sigma=0.01, device=0) as p:
        cost, H, g = p.eval(x0)
        x, s, tr = p.solve(x0)
        Hi, bi, chi, sv = p.information(x)
        T, _, _, _ = p.closed_form()
    scale = np.abs(H).max()
    assert abs(r[0]["cost"] - cost) <= 1e-12 * cost
    np.testing.assert_allclose(r[0]["H"], H, rtol=0, atol=1e-12 * scale)
    np.testing.assert_allclose(r[0]["g"], g, rtol=0, atol=1e-12 * scale)
    ang, dt = oracle.pose_error(r[0]["x"], x)
    assert ang < 1e-9 and dt < 1e-9 and int(r[0]["term"]) == s.termination and int(r[0]["iters"]) == s.num_iterations
    np.testing.assert_allclose(r[0]["costs"], [t.cost for t in tr], rtol=1e-10)
    np.testing.assert_allclose(r[0]["Hi"], Hi, rtol=0, atol=1e-11 * np.abs(Hi).max())
    np.testing.assert_allclose(r[0]["T"], T, atol=1e-9)

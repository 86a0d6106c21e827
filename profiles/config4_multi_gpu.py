"""BASELINE configs[3]: 10^6 frames x 2*10^3 points (2*10^9 residuals, 48 GB) sharded by frame over the GPUs of one box,
full LM to convergence with the all-reduce fused into the sweep kernel.  Launch:
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/config4_multi_gpu.py
(also runs on fewer GPUs, or on one: 48 GB fit in a single B200's 180 GB)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from camlasercalibratool_b200 import Comm, Problem, comm_unique_id, shard_range  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
N, M = int(os.environ.get("C4_FRAMES", 1_000_000)), 2_000
comm = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    comm = Comm(uid[0], world, rank, device=local)

    def all_gather(blob):
        out = [None] * world
        dist.all_gather_object(out, blob)
        return out

    comm.enable_p2p(all_gather)
b, e = shard_range(N, world, rank)
t0 = time.perf_counter()
x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
gt = np.array([0.2, 0.3, -0.1, 0.5, -0.5, 0.5, 0.5])
out = {}
for sigma in (0.0, 0.01):
    with Problem.synthetic(N, M, seed=4, sigma=sigma, frame_begin=b, frame_end=e, device=local) as p:
        torch.cuda.synchronize()
        gen_s = time.perf_counter() - t0
        p.attach_comm(comm)
        p.solve(x0)
        x, s, tr = p.solve(x0)
        k = p.bench_eval(x, 5, flush_l2=False)
        out[f"sigma_{sigma}"] = {"lm_ms": s.device_ms, "sweeps": s.num_sweeps, "lm_iterations": s.num_iterations - 1,
                                 "termination": int(s.termination), "final_cost": s.final_cost,
                                 "residual_evals_per_s": N * M * s.num_sweeps / (s.device_ms * 1e-3),
                                 "lm_iters_per_s": (s.num_iterations - 1) / (s.device_ms * 1e-3),
                                 "translation_error_m": float(np.linalg.norm(x[:3] - gt[:3])),
                                 "quaternion_error": float(min(np.linalg.norm(x[3:] - gt[3:]), np.linalg.norm(x[3:] + gt[3:]))),
                                 "local_sweep_ms": float(np.mean(k)), "local_GBps": p.algorithmic_bytes() / float(np.mean(k)) / 1e6}
    t0 = time.perf_counter()
if rank == 0:
    print(json.dumps({"workload": f"config 4: {N} frames x {M} points over {world} GPU(s), {e - b} frames on rank 0", **out}))
if comm is not None:
    comm.close()
    dist.destroy_process_group()

"""Where the end-to-end time of one CamLaserCalibration-sized call goes (config 2: 10^4 x 10^3 points, 240 MB of host AoS).
Run on the GPU box: python profiles/e2e_breakdown.py  -> JSON on stdout."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from camlasercalibratool_b200 import Problem  # noqa: E402
from camlasercalibratool_b200.api import default_options, pinned_array  # noqa: E402

N, M = 10000, 1000
with Problem.synthetic(N, M, seed=1, sigma=0.01) as g:
    d = g.download()
pin = pinned_array(d["points"].shape)
pin.array[...] = d["points"]
fp, off = d["frame_pose"], d["offsets"]
X0 = np.array([0.05, -0.02, 0.1, 0.02, 0.01, -0.015, 1.0])
X0[3:] /= np.linalg.norm(X0[3:])


def wall(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t))
    return float(np.median(ts)), r


out = {}
host_t = torch.from_numpy(pin.array)
dev_t = torch.empty_like(host_t, device="cuda")
try:
    torch.cuda.cudart().cudaHostRegister  # noqa: B018
except Exception:
    pass
# raw PCIe: the library's own pinned block -> device, one cudaMemcpyAsync
out["raw_h2d_ms"], _ = wall(lambda: dev_t.copy_(host_t, non_blocking=True))
out["raw_h2d_gbs"] = pin.nbytes / out["raw_h2d_ms"] / 1e6
holder = {}


def create():
    if "q" in holder:
        holder["q"].close()
    holder["q"] = Problem.from_arrays(fp, off, pin.array)


out["create_ms"], _ = wall(create)
q = holder["q"]
out["solve_ms"], r = wall(lambda: q.solve(X0, default_options()))
out["solve_sweeps"] = r[1].num_sweeps


def full():
    with Problem.from_arrays(fp, off, pin.array) as qq:
        return qq.solve(X0, default_options())


out["create_solve_destroy_ms"], _ = wall(full)
q.close()
print(json.dumps(out))

"""Device time of full LM solves under the three loop drivers (CLC_LOOP_IN_KERNEL = 0 launch per iteration, 1 single-block
problems loop in the kernel, 2 persistent grid for every problem), at the reference's size and at BASELINE configs[1].
    python profiles/loop_mode_timing.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200 import Problem  # noqa: E402

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
for name, frames, beams, planar in (("config1 50x180", 50, 180, 0), ("config2 10^4x10^3 general", 10000, 1000, 0),
                                    ("config2 10^4x10^3 planar", 10000, 1000, 1), ("config3 10^5x2000 planar", 100000, 2000, 1)):
    for mode in (0, 1, 2):
        os.environ["CLC_LOOP_IN_KERNEL"] = str(mode)
        with Problem.synthetic(frames, beams, seed=7, sigma=0.01) as p:
            p.set_planar_mode(planar)
            for _ in range(3):
                p.solve(X0)
            ms, sw = [], 0
            for _ in range(20 if frames <= 10000 else 3):
                x, s, _ = p.solve(X0)
                ms.append(s.device_ms)
                sw = s.num_sweeps
            print(f"{name:28s} mode {mode}: device ms median {np.median(ms):8.4f} min {np.min(ms):8.4f}  sweeps {sw}  "
                  f"us/sweep {1e3 * np.median(ms) / sw:7.2f}  planar={p.planar}")

# ---- host polls: LM iterations enqueued between two polls of the device `done` flag (clc_lm_options.iterations_per_sync) ----
from camlasercalibratool_b200 import default_options  # noqa: E402

os.environ["CLC_LOOP_IN_KERNEL"] = "1"
for name, frames, beams, planar in (("config2 general", 10000, 1000, 0), ("config2 planar", 10000, 1000, 1)):
    with Problem.synthetic(frames, beams, seed=7, sigma=0.01) as p:
        p.set_planar_mode(planar)
        for ips in (4, 8, 16, 32):
            opt = default_options(iterations_per_sync=ips)
            for _ in range(3):
                p.solve(X0, opt)
            ms = []
            for _ in range(20):
                x, s, _ = p.solve(X0, opt)
                ms.append(s.device_ms)
            print(f"{name:18s} iterations_per_sync {ips:3d}: device ms median {np.median(ms):8.4f} min {np.min(ms):8.4f}  sweeps {s.num_sweeps}")
        gt_start = x
        for ips in (8, 16):
            opt = default_options(iterations_per_sync=ips)
            ms = []
            for _ in range(20):
                _, s, _ = p.solve(gt_start, opt)
                ms.append(s.device_ms)
            print(f"{name:18s} start at the solution, iterations_per_sync {ips:3d}: device ms median {np.median(ms):8.4f}  sweeps {s.num_sweeps}")

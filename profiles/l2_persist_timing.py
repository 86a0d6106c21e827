"""L2 residency across LM iterations (CLC_L2_PERSIST_MB: persisting access-policy window over the x,y block during solves):
device time of full LM solves at BASELINE configs[1] for several set-aside sizes, planar and general kernels, and at configs[2].
    python profiles/l2_persist_timing.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200 import Problem  # noqa: E402

X0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
for name, frames, beams, planar, reps in (("config2 planar (160 MB)", 10000, 1000, 1, 20), ("config2 general (240 MB)", 10000, 1000, 0, 20),
                                          ("5000x1000 planar (80 MB)", 5000, 1000, 1, 20), ("config3 planar (3.2 GB)", 100000, 2000, 1, 3)):
    for rep in range(2):
        for mb in (0, 16, 32, 48, 64, 96, 120):
            os.environ["CLC_L2_PERSIST_MB"] = str(mb)
            with Problem.synthetic(frames, beams, seed=7, sigma=0.01) as p:
                p.set_planar_mode(planar)
                for _ in range(3):
                    p.solve(X0)
                ms = []
                for _ in range(reps):
                    x, s, _ = p.solve(X0)
                    ms.append(s.device_ms)
                k = p.bench_eval(x, 20, True)
                print(f"{name:26s} pass {rep} persist {mb:4d} MB: solve ms median {np.median(ms):8.4f} min {np.min(ms):8.4f}  sweeps {s.num_sweeps}  "
                      f"us/sweep {1e3 * np.median(ms) / s.num_sweeps:7.2f}   flushed single launch {1e3 * k.mean():7.2f} us", flush=True)

"""Per-launch time of the sweep kernel, general (24 B/point) vs planar (16 B/point), L2 flushed between launches.
python profiles/planar_kernel_timing.py  -> one JSON line (run on the GPU box)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from camlasercalibratool_b200 import Problem  # noqa: E402

X = np.array([0.05, -0.02, 0.1, 0, 0, 0, 1.0])
out = {}
for name, (n, m) in {"config2": (10_000, 1_000), "config3": (100_000, 2_000)}.items():
    with Problem.synthetic(n, m, seed=1, sigma=0.01) as g:
        row = {}
        for mode, label in ((0, "general"), (1, "planar")):
            g.set_planar_mode(mode)
            g.bench_eval(X, 5, flush_l2=True)
            ms = g.bench_eval(X, 40, flush_l2=True)
            b = g.streamed_bytes()
            row[label] = {"us_mean": 1e3 * float(np.mean(ms)), "us_min": 1e3 * float(np.min(ms)), "bytes": b,
                          "GBps": b / float(np.mean(ms)) / 1e6}
            g.solve(np.array([0, 0, 0, 0, 0, 0, 1.0]))
            sm = []
            for _ in range(6):
                _, s, _ = g.solve(np.array([0, 0, 0, 0, 0, 0, 1.0]))
                sm.append(s.device_ms)
            row[label]["solve_ms"] = float(np.median(sm))
            row[label]["solve_ms_all"] = sm
            row[label]["sweeps"] = s.num_sweeps
        out[name] = row
print(json.dumps(out))

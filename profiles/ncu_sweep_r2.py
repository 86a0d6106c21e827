"""Target of the round-2 ncu --set full capture: L2-flushed launches of the sweep kernel at
  config 2 (10^4 x 10^3): planar x2, general x2;  config 3 (10^5 x 2000): planar x1, general x1;
  config 5 (10^5 x 2000 + 2*10^5 board-edge residuals, equidistant camera chain): general x1.
ncu --set full --clock-control none --import-source on -k regex:clc_sweep -c 8 -o gpurun_out/r2_sweep python profiles/ncu_sweep_r2.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from camlasercalibratool_b200 import Problem  # noqa: E402

X = np.array([0.05, -0.02, 0.1, 0, 0, 0, 1.0])
with Problem.synthetic(10_000, 1_000, seed=1, sigma=0.01) as g:
    assert g.planar
    g.bench_eval(X, 2, flush_l2=True)
    g.set_planar_mode(0)
    g.bench_eval(X, 2, flush_l2=True)
with Problem.synthetic(100_000, 2_000, seed=1, sigma=0.01) as g:
    g.bench_eval(X, 1, flush_l2=True)
    g.set_planar_mode(0)
    g.bench_eval(X, 1, flush_l2=True)
with Problem.synthetic(100_000, 2_000, seed=1, sigma=0.01, with_edges=True, camera="equi", pixel_sigma=0.3) as g:
    g.set_planar_mode(0)
    g.bench_eval(X, 1, flush_l2=True)
print("done")

"""Target of the ncu --set full capture: config 2, three L2-flushed launches of the planar (two-stream) sweep kernel, then
three of the general (three-stream) one.
ncu --set full --clock-control none --import-source on -k regex:clc_sweep -c 6 -o gpurun_out/sweep_families python profiles/ncu_sweep_families.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from camlasercalibratool_b200 import Problem  # noqa: E402

X = np.array([0.05, -0.02, 0.1, 0, 0, 0, 1.0])
with Problem.synthetic(10_000, 1_000, seed=1, sigma=0.01) as g:
    assert g.planar
    g.bench_eval(X, 3, flush_l2=True)
    g.set_planar_mode(0)
    g.bench_eval(X, 3, flush_l2=True)
print("done")

"""Where do the ~3.5 us of the on-device lm_update go?  Needs the experiment build gpurun_variants/libclc_lmprof.so
(-DCLC_LM_PROFILE: clock64 stamps at the section boundaries of the last update).
    CLC_LIB_PATH=gpurun_variants/libclc_lmprof.so python profiles/lm_update_profile.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200 import Problem, _lib, default_options  # noqa: E402

L = _lib.load()
L.clc_debug_lm_profile.argtypes = [C.POINTER(C.c_longlong)]
names = ["entry -> accept/reject decided (gradient_max_norm incl.)", "-> iteration recorded, termination tests", "-> scaled damped system built",
         "-> Cholesky 6x6 solve", "-> model cost change", "-> candidate pose (pose_plus)"]
x0 = np.array([0, 0, 0, 0, 0, 0, 1.0])
os.environ["CLC_LOOP_IN_KERNEL"] = "0"
with Problem.synthetic(10000, 1000, seed=7, sigma=0.01) as p:
    for cap in (3, 6):  # stop after an accepted step: the stamps are those of a full update
        p.solve(x0, default_options(max_num_iterations=cap))
        buf = (C.c_longlong * 16)()
        L.clc_debug_lm_profile(buf)
        t = np.array(buf[:7], dtype=np.int64)
        d = np.diff(t)
        print(f"after {cap} iterations: total {t[6] - t[0]} cycles = {(t[6] - t[0]) / 1.965e3:.2f} us at 1965 MHz")
        for n, v in zip(names, d):
            print(f"   {n:62s} {v:6d} cycles  {v / 1.965e3:5.2f} us")

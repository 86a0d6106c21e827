"""A/B of library variants on ONE box: per-launch time of the sweep kernel (CUDA events, L2 flushed) at BASELINE configs[1] and
configs[2], general and planar families.  Minimal ctypes surface so that older builds of the library load too.
    python profiles/variant_ab.py lib1.so lib2.so ..."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200._lib import SyntheticDesc  # noqa: E402  (plain ctypes struct, no library needed)


def run(path, frames, beams, planar, n):
    L = C.CDLL(path)
    L.clc_last_error.restype = C.c_char_p
    d = SyntheticDesc()
    d.n_frames_total, d.frame_begin, d.frame_end, d.beams, d.seed, d.sigma = frames, 0, frames, beams, 7, 0.01
    d.with_edges, d.use_loss, d.cauchy_a, d.device = 0, 1, 0.05, 0
    h = C.c_void_p()
    assert L.clc_problem_create_synthetic(C.byref(h), C.byref(d)) == 0, L.clc_last_error()
    assert L.clc_problem_set_planar_mode(h, planar) == 0
    x = (C.c_double * 7)(0, 0, 0, 0, 0, 0, 1.0)
    ms = (C.c_float * n)()
    L.clc_bench_eval(h, x, 5, 1, ms)
    out = []
    for _ in range(3):
        assert L.clc_bench_eval(h, x, n, 1, ms) == 0, L.clc_last_error()
        out.append(np.array(ms[:]))
    L.clc_problem_destroy(h)
    a = np.concatenate(out) * 1e3
    return a


if __name__ == "__main__":
    libs = sys.argv[1:]
    for frames, beams, n in ((10000, 1000, 100), (100000, 2000, 20)):
        for planar in (0, 1):
            for rep in range(2):  # interleaved repetitions: drift of the box shows up as a difference between the two passes
                for lib in libs:
                    a = run(os.path.abspath(lib), frames, beams, planar, n)
                    print(f"{frames}x{beams} planar={planar} pass {rep} {os.path.basename(lib):22s} mean {a.mean():8.2f} us  median {np.median(a):8.2f}  min {a.min():8.2f}  p90 {np.percentile(a, 90):8.2f}",
                          flush=True)

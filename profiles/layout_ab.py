"""Does the relative placement of the coordinate arrays matter?  Per-launch time of the sweep kernel at BASELINE configs[2]
(10^5 x 2000) for several (CLC_SKEW_Y, CLC_SKEW_Z) placements, next to an older build of the library.
    python profiles/layout_ab.py [old_lib.so]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np  # noqa: E402
from variant_ab import run  # noqa: E402

cur = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "camlasercalibratool_b200", "libclc_b200.so")
old = sys.argv[1] if len(sys.argv) > 1 else None
pairs = [(None, None), (0, 0), (124928 + 1048576, 1048576), (4096, 8192)]  # None: the library's default placement
os.environ["CLC_DEBUG_LAYOUT"] = "1"
for rep in range(3):
    for planar in (0, 1):
        if old:
            a = run(os.path.abspath(old), 100000, 2000, planar, 20)
            print(f"pass {rep} planar={planar} OLD lib                     mean {a.mean():8.2f} us  median {np.median(a):8.2f}  min {a.min():8.2f}", flush=True)
        for sy, sz in pairs:
            for k, v in (("CLC_SKEW_Y", sy), ("CLC_SKEW_Z", sz)):
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = str(v)
            a = run(cur, 100000, 2000, planar, 20)
            print(f"pass {rep} planar={planar} skew_y {str(sy):>8s} skew_z {str(sz):>8s}   mean {a.mean():8.2f} us  median {np.median(a):8.2f}  min {a.min():8.2f}", flush=True)

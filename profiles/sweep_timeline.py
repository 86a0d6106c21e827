"""Per-block timeline of one sweep from the library's profiling hook (clc_debug_sweep_timing): where the time of the
~65 us kernel goes (ramp, streaming, per-warp tail, block reduce, final reduce, LM update).  Run on the GPU box:
    python profiles/sweep_timeline.py [frames] [beams]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from camlasercalibratool_b200 import Problem, _lib  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
beams = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
L = _lib.load()
L.clc_debug_sweep_timing.argtypes = [C.c_void_p, _lib.c_double_p, C.c_int, C.c_int, C.POINTER(C.c_ulonglong), C.POINTER(C.c_int),
                                     C.POINTER(C.c_ulonglong)]
planar_mode = int(os.environ.get("TIMELINE_PLANAR", "-1"))
x = np.array([0, 0, 0, 0, 0, 0, 1.0])
with Problem.synthetic(frames, beams, seed=7, sigma=0.01) as p:
    if planar_mode >= 0:
        p.set_planar_mode(planar_mode)
    print(f"frames={frames} beams={beams} planar={p.planar}")
    p.bench_eval(x, 5, True)
    for with_lm in (0, 1):
        for flush in (1, 0):
            rows, keep, wkeep = [], [], []
            for rep in range(7):
                buf = (C.c_ulonglong * (8 * 4096))()
                wbuf = (C.c_ulonglong * (16 * 4096))()
                grid = C.c_int()
                _lib.check(L.clc_debug_sweep_timing(p._h, x.ctypes.data_as(_lib.c_double_p), with_lm, flush, buf, C.byref(grid), wbuf), "timing")
                t = np.array(buf[: 8 * grid.value], dtype=np.float64).reshape(grid.value, 8)
                t0 = t[:, 0].min()
                keep.append(t.copy())
                wkeep.append(np.array(wbuf[: 16 * grid.value], dtype=np.float64).reshape(grid.value, 16) - t0)
                last = int(np.argmax(t[:, 4]))
                rows.append([t[:, 0].max() - t0, np.median(t[:, 1]) - t0, t[:, 1].max() - t0, t[:, 2].max() - t0, t[:, 3].max() - t0,
                             t[last, 4] - t0, t[last, 5] - t0])
            if with_lm == 0 and flush == 1:
                # per-warp "stream done": is the straggler tail a property of the SM (same blocks late in every run) or noise?
                w = np.array(wkeep) / 1e3  # [rep, block, warp] us
                print(f"  per-warp stream done (us): median {np.median(w):.2f}  p90 {np.percentile(w, 90):.2f}  p99 {np.percentile(w, 99):.2f}  "
                      f"max {w.max(axis=(1, 2)).mean():.2f} (mean over reps of the slowest warp)  min {w.min(axis=(1, 2)).mean():.2f}")
                bm = w.mean(axis=2)  # [rep, block]
                print(f"  block means: spread over blocks {bm.std(axis=1).mean():.2f} us, within-block spread {w.std(axis=2).mean():.2f} us, "
                      f"rep-to-rep correlation of block means {np.corrcoef(bm)[0, 1:].mean():.2f}")
                sm = np.array(keep)[0][:, 7].astype(int)
                late = np.argsort(-bm.mean(axis=0))[:8]
                print("  latest blocks (block:smid:mean us): " + " ".join(f"{b}:{sm[b]}:{bm[:, b].mean():.1f}" for b in late))
                wm = w.mean(axis=(0, 1))
                print("  mean by warp index: " + " ".join(f"{v:.1f}" for v in wm))
                np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "timeline_warps.npy"), w)
                np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "timeline_blocks.npy"), np.array(keep))
            r = np.median(np.array(rows), axis=0) / 1e3
            print(f"lm={with_lm} flush_l2={flush} grid={grid.value}: last block start {r[0]:6.2f} | stream done median {r[1]:6.2f} max {r[2]:6.2f} | "
                  f"tiles flushed {r[3]:6.2f} | partials written {r[4]:6.2f} | final sums {r[5]:6.2f} | after LM {r[6]:6.2f}  (us since first block start)")

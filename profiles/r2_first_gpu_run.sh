#!/bin/bash
# round 2, first GPU call: tests, the C++ end-to-end driver with a pack-thread sweep, a short bench
set -x
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|^CPU\(s\)|NUMA|Socket" ; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for t in 1 2 4 8 16 32; do
  CLC_PACK_THREADS=$t timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 5 2 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/threads=$t /"
done | tee gpurun_out/r2_dropin_thread_sweep.txt
for c in 65536 131072 524288 1048576; do
  CLC_UPLOAD_CHUNK_POINTS=$c timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 5 2 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/chunk=$c /"
done | tee gpurun_out/r2_dropin_chunk_sweep.txt
timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 20 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | tee gpurun_out/r2_dropin_config1.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-config3 2>&1 | tail -3 | tee gpurun_out/r2_bench_first.json
timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -6 | tee gpurun_out/r2_timeline_first.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2_carveout_experiment.txt
import os, numpy as np
from camlasercalibratool_b200 import Problem
x = np.array([0, 0, 0, 0, 0, 0, 1.0])
with Problem.synthetic(10000, 1000, seed=7, sigma=0.01) as p:
    p.set_planar_mode(0)
    for mode in ("1", "0", "1", "0"):
        os.environ["CLC_FLUSH_SMEM"] = mode
        p.bench_eval(x, 5, True)
        ms = p.bench_eval(x, 100, True)
        print(f"general  CLC_FLUSH_SMEM={mode}: mean {ms.mean()*1e3:.2f} us  min {ms.min()*1e3:.2f}  median {np.median(ms)*1e3:.2f}")
    p.set_planar_mode(1)
    for mode in ("1", "0"):
        os.environ["CLC_FLUSH_SMEM"] = mode
        p.bench_eval(x, 5, True)
        ms = p.bench_eval(x, 100, True)
        print(f"planar   CLC_FLUSH_SMEM={mode}: mean {ms.mean()*1e3:.2f} us  min {ms.min()*1e3:.2f}  median {np.median(ms)*1e3:.2f}")
PY

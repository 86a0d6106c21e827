#!/bin/bash
# round 2, second GPU call: tests on the new block reduction, new bench.py (C++ drop-in e2e, config 3/5 rows, strong leg, check block), per-warp timeline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
CLC_DROPIN_TIMING=1 CLC_UPLOAD_TIMING=1 timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 5 2 2>&1 | grep -E "CLC_DROPIN|CLC_UPLOAD|rror" | tail -12 | tee gpurun_out/r2_dropin_phases.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 3000 gpurun_out/r2_bench_n1.err; cat gpurun_out/r2_bench_n1.json | cut -c1-6000
TIMELINE_PLANAR=0 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -16 | tee gpurun_out/r2_timeline_general.txt
TIMELINE_PLANAR=1 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -16 | tee gpurun_out/r2_timeline_planar.txt

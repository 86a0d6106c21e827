#!/bin/bash
# round 2, third GPU call: tests, loop drivers, NUMA placement of the pack threads, bench, timelines
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python profiles/loop_mode_timing.py 2>&1 | tee gpurun_out/r2_loop_modes.txt
for aff in "" "taskset -c 0-31,64-95" "taskset -c 32-63,96-127" "taskset -c 0-15" "taskset -c 32-47"; do
  CLC_DROPIN_TIMING=1 timeout 300 $aff camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 5 2 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[$aff] /" | cut -c1-700
done | tee gpurun_out/r2_dropin_affinity.txt
nvidia-smi topo -m 2>&1 | head -20 | tee gpurun_out/r2_topo.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 2000 gpurun_out/r2_bench_n1.err; cut -c1-3000 gpurun_out/r2_bench_n1.json
TIMELINE_PLANAR=0 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -16 | tee gpurun_out/r2_timeline_general.txt
TIMELINE_PLANAR=1 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -16 | tee gpurun_out/r2_timeline_planar.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; tail -c 800 gpurun_out/r2_bench_final_n1.err; cut -c1-600 gpurun_out/r2_bench_final_n1.json
TIMELINE_PLANAR=0 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -9 | tee gpurun_out/r2_timeline_config2_general.txt
TIMELINE_PLANAR=1 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -9 | tee gpurun_out/r2_timeline_config2_planar.txt

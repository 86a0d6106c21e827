#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
CLC_LIB_PATH=gpurun_variants/libclc_lmprof.so timeout 300 python profiles/lm_update_profile.py 2>&1 | tail -20 | tee gpurun_out/r2_lm_update_profile.txt
timeout 300 python profiles/loop_mode_timing.py 2>&1 | tail -26 | tee gpurun_out/r2_loop_modes3.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 1500 gpurun_out/r2_bench_n1.err; cut -c1-1500 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-600

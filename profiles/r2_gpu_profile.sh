#!/bin/bash
# round 2: ncu captures (never a bench number): --set full of the sweep kernel families, and the launch list of a short bench
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:clc_sweep -c 8 -f -o gpurun_out/r2_sweep python profiles/ncu_sweep_r2.py 2>&1 | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --kernel-launches 5 --no-cpu-baseline --no-strong > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2_launches.csv | cut -c1-300
timeout 300 python profiles/loop_mode_timing.py 2>&1 | tail -14 | tee gpurun_out/r2_loop_modes2.txt
timeout 300 python profiles/widened_rows_timing.py 2>&1 | tail -5 | tee gpurun_out/r2_widened_rows.txt
CLC_LIB_PATH=gpurun_variants/libclc_lmprof.so timeout 300 python profiles/lm_update_profile.py 2>&1 | tail -20 | tee gpurun_out/r2_lm_update_profile.txt

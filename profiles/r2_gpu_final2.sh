#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; tail -c 500 gpurun_out/r2_bench_final_n1.err; cut -c1-300 gpurun_out/r2_bench_final_n1.json
timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 20 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-600 | tee gpurun_out/r2_dropin_config1_final.txt

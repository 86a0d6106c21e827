#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_loop_modes.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_cpp_dropin.py tests/test_gpu_degenerate.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python profiles/loop_mode_timing.py 2>&1 | head -4 | tee gpurun_out/r2_loop_modes4.txt
timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 20 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-700 | tee gpurun_out/r2_dropin_config1_small.txt

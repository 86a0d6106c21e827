#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_loop_modes.py tests/test_gpu_golden.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 200 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 20 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-330

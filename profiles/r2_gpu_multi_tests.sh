#!/bin/bash
set -x
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gather_group.py -m gpu -q 2>&1 | tail -6
CLC_DEVICES=0,1 timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 10 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-600
CLC_DEVICES=0,1 timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 20000 1000 0.01 7 5 2 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-600

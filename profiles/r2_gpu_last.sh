#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_gather_group.py tests/test_cpp_dropin.py tests/test_upload_pack.py -m gpu -q 2>&1 | tail -3
timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-420

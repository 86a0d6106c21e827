#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 50 180 0.01 1 20 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | cut -c1-700 | tee gpurun_out/r2_dropin_config1_small2.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

#!/bin/bash
# round 2, fifth GPU call: placement of the coordinate arrays (config 3), then the GPU tests on the new allocation code
set -x
mkdir -p gpurun_out
timeout 900 python profiles/layout_ab.py gpurun_variants/libclc_old.so 2>&1 | tee gpurun_out/r2_layout_ab.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python profiles/layout_ab.py gpurun_variants/libclc_old.so 2>&1 | grep -v DEBUG_LAYOUT | tee gpurun_out/r2_layout_ab2.txt
timeout 600 python profiles/variant_ab.py gpurun_variants/libclc_old.so camlasercalibratool_b200/libclc_b200.so 2>&1 | tee gpurun_out/r2_variant_ab2.txt

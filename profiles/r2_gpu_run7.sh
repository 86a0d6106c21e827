#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python profiles/variant_ab.py gpurun_variants/libclc_old.so camlasercalibratool_b200/libclc_b200.so gpurun_variants/libclc_fixed.so gpurun_variants/libclc_old.so gpurun_variants/libclc_fixed.so camlasercalibratool_b200/libclc_b200.so 2>&1 | tee gpurun_out/r2_variant_ab3.txt

#!/bin/bash
# round 2, fourth GPU call: A/B of kernel variants on one box, config-1 timeline, write-combined upload slots
set -x
mkdir -p gpurun_out
timeout 900 python profiles/variant_ab.py gpurun_variants/libclc_old.so camlasercalibratool_b200/libclc_b200.so gpurun_variants/libclc_ls1.so gpurun_variants/libclc_ls2.so gpurun_variants/libclc_ls3.so 2>&1 | tee gpurun_out/r2_variant_ab.txt
TIMELINE_PLANAR=0 timeout 300 python profiles/sweep_timeline.py 50 180 2>&1 | tail -16 | tee gpurun_out/r2_timeline_config1.txt
for wc in 0 1; do
  CLC_UPLOAD_WC=$wc timeout 300 taskset -c 0-31,64-95 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[wc=$wc node0] /" | cut -c1-520
  CLC_UPLOAD_WC=$wc timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[wc=$wc free] /" | cut -c1-520
done | tee gpurun_out/r2_dropin_wc.txt
for ls in ls1 ls2 ls3; do
  CLC_LIB_PATH=gpurun_variants/libclc_$ls.so TIMELINE_PLANAR=0 timeout 300 python profiles/sweep_timeline.py 2>&1 | tail -16 | head -6 | sed "s/^/[$ls] /"
done | tee gpurun_out/r2_timeline_lockstep.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
for c in 65536 131072 262144 524288 1048576 2097152; do
  CLC_UPLOAD_CHUNK_POINTS=$c timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[chunk=$c] /" | cut -c1-420
done | tee gpurun_out/r2_dropin_chunk_sweep2.txt
for k in 3 4 6 8; do
  CLC_UPLOAD_SLOTS=$k timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[slots=$k] /" | cut -c1-420
done | tee -a gpurun_out/r2_dropin_chunk_sweep2.txt
CLC_UPLOAD_SLOTS=8 CLC_UPLOAD_CHUNK_POINTS=524288 timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[slots=8 chunk=512K] /" | cut -c1-420 | tee -a gpurun_out/r2_dropin_chunk_sweep2.txt
CLC_PACK_THREADS=12 timeout 300 camlasercalibratool_b200/host/clc_dropin_bench 10000 1000 0.01 7 8 3 2>&1 | grep -E "CLC_DROPIN_JSON|rror" | sed "s/^/[threads=12] /" | cut -c1-420 | tee -a gpurun_out/r2_dropin_chunk_sweep2.txt
timeout 600 python -m pytest tests/test_gpu_loop_modes.py -m gpu -x -q 2>&1 | tail -3

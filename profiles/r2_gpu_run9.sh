#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python profiles/l2_persist_timing.py 2>&1 | tee gpurun_out/r2_l2_persist.txt

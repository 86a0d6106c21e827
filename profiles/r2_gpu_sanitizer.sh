#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python profiles/sanitizer_smoke.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_memcheck.txt
timeout 800 compute-sanitizer --tool racecheck --error-exitcode 7 python profiles/sanitizer_smoke.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/r2_sanitizer_racecheck.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-strong --no-config3 --no-cpu-baseline > gpurun_out/r2_bench_clock_check.json 2> gpurun_out/r2_bench_clock_check.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_clock_check.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_clock_check.json'))
print("value %.4g clocks %s e2e %.3f"%(d['value'], d['clocks'], d['e2e']['ms_per_step']))
PY

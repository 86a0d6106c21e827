#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-strong --no-config3 --no-cpu-baseline > gpurun_out/r2_bench_verify.json 2> gpurun_out/r2_bench_verify.err; tail -c 500 gpurun_out/r2_bench_verify.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_verify.json'))
print("value", d['value'], "clocks", d['clocks'], "e2e", d['e2e']['ms_per_step'], d['e2e']['ms_per_step_max'], d['e2e']['phases_ms'], d['e2e']['pack_threads'], "c1", d['e2e']['config1']['body_ms'])
PY

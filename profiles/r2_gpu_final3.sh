#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final3_n1.json 2> gpurun_out/r2_bench_final3_n1.err; echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_final3_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_final3_n1.json'))
print("value %.4g ms %.4f frac %.3f planar %.3f c3 %.3f c5 %.3f e2e %.3f ms upload %.3f c1 %.3f clocks %s"%(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['planar']['frac'], d['roofline']['config3']['frac'], d['roofline']['config5']['frac'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']['upload'], d['e2e']['config1']['body_ms'], d['clocks']))
PY

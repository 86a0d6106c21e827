#!/bin/bash
# what the driver runs at round end, on the committed HEAD: GPU tests, smoke, the N=1 bench line, the reference arm
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_head_n1.json 2> gpurun_out/r2_bench_head_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_head_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_head_n1.json'))
print("value %.4g ms %.4f frac %.3f planar %.3f c3 %.3f c5 %.3f e2e %.3f ms (max %.3f) upload %.3f c1 %.3f clocks %s launches %d"%(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['planar']['frac'], d['roofline']['config3']['frac'], d['roofline']['config5']['frac'], d['e2e']['ms_per_step'], d['e2e']['ms_per_step_max'], d['e2e']['phases_ms']['upload'], d['e2e']['config1']['body_ms'], d['clocks'], d['gpu_launches']))
PY

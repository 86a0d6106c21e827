#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n2.json 2> gpurun_out/r2_bench_final_n2.err
tail -c 600 gpurun_out/r2_bench_final_n2.err
python - <<'PY'
import json
txt=open('gpurun_out/r2_bench_final_n2.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print("value", d['value'], d['ms_per_step'], d['config']['check'], d['clocks'], d['e2e']['ms_per_step'], d['e2e']['phases_ms'], d['config']['strong_scaling']['general_24B']['ms_per_solve'])
PY

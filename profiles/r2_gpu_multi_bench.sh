#!/bin/bash
# multi-GPU bench lines only (gpurun --gpus 8): N = 1, 2, 4, 8 back to back on one box, as the driver's scaling run does
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-config3 --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; cut -c1-400 gpurun_out/r2_scale_n1.json
for n in 2 4 8; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  tail -c 600 gpurun_out/r2_scale_n$n.err; grep '^{' gpurun_out/r2_scale_n$n.json | cut -c1-400
done

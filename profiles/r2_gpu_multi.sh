#!/bin/bash
# round 2, multi-GPU call (gpurun --gpus N): N>1 parity tests, the weak/strong bench line at N ranks, the in-process group
# through the C++ drop-in.  usage: bash profiles/r2_gpu_multi.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_gather_group.py -m gpu -x -q 2>&1 | tail -8
for n in $(seq 2 $N); do
  case $n in 2|4|8) ;; *) continue;; esac
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  tail -c 1500 gpurun_out/r2_bench_n$n.err; cut -c1-2500 gpurun_out/r2_bench_n$n.json
done

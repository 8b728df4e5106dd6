#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_8gpu_c41.json; cut -c1-260 gpurun_out/bench_8gpu_c41.json

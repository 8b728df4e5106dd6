#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_4gpu_c69.json; cut -c1-300 gpurun_out/bench_4gpu_c69.json

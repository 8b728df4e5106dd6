#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graded_gpu.py -q -m gpu -k "two_gpu" 2>&1 | tail -4 | tee gpurun_out/pytest_2gpu_c61.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_2gpu_c61.json; cut -c1-260 gpurun_out/bench_2gpu_c61.json

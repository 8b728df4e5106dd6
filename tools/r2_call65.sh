#!/usr/bin/env bash
# validation of the final binary (after the DFT row padding and the attention load pipelining)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_c65.log; tail -3 gpurun_out/pytest_gpu_c65.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_c65.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c65.json; cut -c1-200 gpurun_out/bench_c65.json

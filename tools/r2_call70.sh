#!/usr/bin/env bash
# validation of the binary with the GroupNorm sums taken inside combine_kernel
mkdir -p gpurun_out
timeout 150 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_c70.log; tail -3 gpurun_out/pytest_gpu_c70.log
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_c70.log

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_c47.log; tail -3 gpurun_out/pytest_gpu_c47.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_c47.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c47.json; cut -c1-200 gpurun_out/bench_c47.json
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c47.csv python tools/profile_eval.py | tail -1

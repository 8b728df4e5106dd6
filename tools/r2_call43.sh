#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "in_kernel_statistics or gn_act" 2>&1 | grep -E "^E  |passed|failed|Error" | head
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_c43.log; tail -3 gpurun_out/pytest_gpu_c43.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_c43.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c43.json; cut -c1-200 gpurun_out/bench_c43.json
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c43.csv python tools/profile_eval.py | tail -1

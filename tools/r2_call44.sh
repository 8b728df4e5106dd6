#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "in_kernel_statistics or gn_act" 2>&1 | grep -E "^E  |passed|failed|Error" | head
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c44.csv python tools/profile_eval.py | tail -1

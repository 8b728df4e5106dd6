#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k attention 2>&1 | tail -3
DSEP_CUDA_GRAPH=0 bash tools/ncu_membound.sh
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c11.csv python tools/profile_eval.py | tail -1

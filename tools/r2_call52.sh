#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "tap_gather or narrow_conv" 2>&1 | tail -5
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_graded_gpu.py -q -m gpu -x 2>&1 | tail -3
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c52.csv python tools/profile_eval.py | tail -1
for i in 1 2; do
DSEP_PYR_TAPS_MIN=0 timeout 300 python tools/profile_eval.py 2>&1 | tail -1
timeout 300 python tools/profile_eval.py 2>&1 | tail -1
DSEP_PYR_TAPS_MIN=1024 timeout 300 python tools/profile_eval.py 2>&1 | tail -1
done
} > gpurun_out/call52.log 2>&1

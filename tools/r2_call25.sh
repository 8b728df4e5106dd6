#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "im2col or fir_resample_f32" 2>&1 | tail -30 > gpurun_out/pytest_c25.log; tail -30 gpurun_out/pytest_c25.log
DSEP_CONV_IN_COL=0 timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300
DSEP_FIR_F32=0 timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-300

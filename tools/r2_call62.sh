#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graded_gpu.py -q -m gpu -k "two_gpu" > gpurun_out/pytest_2gpu_c62.log 2>&1
DSEP_PYR_TAPS_MIN=0 timeout 600 python -m pytest tests/test_graded_gpu.py -q -m gpu -k "two_gpu" 2>&1 | tail -3 > gpurun_out/pytest_2gpu_c62_notaps.log
DSEP_FIR_KT=1 timeout 600 python -m pytest tests/test_graded_gpu.py -q -m gpu -k "two_gpu" 2>&1 | tail -3 > gpurun_out/pytest_2gpu_c62_kt1.log

#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graded_gpu.py -q -m gpu -k "two_gpu" > gpurun_out/pytest_2gpu_c63.log 2>&1
tail -3 gpurun_out/pytest_2gpu_c63.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "stft or pre_process or post_process or score_model" 2>&1 | tail -3

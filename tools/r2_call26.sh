#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "input_conv_over" 2>&1 | grep -E "^E  |passed|failed|Error" | head -30

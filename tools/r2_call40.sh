#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -x 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c40.json; cut -c1-200 gpurun_out/bench_c40.json

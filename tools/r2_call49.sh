#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-130

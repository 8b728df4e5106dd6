#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "training_forward or compute_score_loss" 2>&1 | grep -E "^E  |passed|failed|Error" | head -20

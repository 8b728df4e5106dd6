#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu_c9.log; tail -6 gpurun_out/pytest_gpu_c9.log
L=gpurun_out/attn_c9.log; : > $L
for tc in 1 0; do for shp in "8 1920 256" "32 256 256" "16 512 256" "32 16 256"; do DSEP_ATTN_TC=$tc timeout 120 python tools/profile_attention.py $shp 2>&1 | tail -1 >> $L; done; done
cat $L
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1

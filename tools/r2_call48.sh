#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
DSEP_PYR_WIDE=1 timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
DSEP_PYR_WIDE=1 timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "nf128 or layerwise" 2>&1 | tail -2
DSEP_PYR_WIDE=1 DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c48.csv python tools/profile_eval.py | tail -1
for i in 1 2; do
timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-120
DSEP_PYR_WIDE=1 timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-120
done

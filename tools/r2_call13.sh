#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c13.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for d in 0 1 2; do run DSEP_CONV_PAIR=1 DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
run DSEP_CONV_PAIR=1 DSEP_REPS=400
run DSEP_CONV_PAIR=1 DSEP_RES=1 DSEP_REPS=20
run DSEP_REPS=20
cat $L
DSEP_CONV_PAIR=1 timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_graded_gpu.py -q -m gpu -k "fused or level0 or e4m3" 2>&1 | tail -3

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "conv" 2>&1 | tail -2
L=gpurun_out/conv_modes_c46.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=20
run DSEP_REPS=400
run DSEP_RES=1 DSEP_REPS=400
run DSEP_SHORT=256 DSEP_REPS=20
cat $L

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c32.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for i in 1 2; do
run DSEP_RES=1 DSEP_REPS=20
run DSEP_RES=1 DSEP_CONV_DEBUG=8 DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=400
run DSEP_RES=1 DSEP_CONV_DEBUG=8 DSEP_REPS=400
done
cat $L

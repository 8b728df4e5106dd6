#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c14.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for d in 0 1 2 4 6; do run DSEP_RES=1 DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
for d in 0 2; do run DSEP_STATS=0 DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
cat $L

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c8.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=400
run DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CIN=256 DSEP_COUT=128 DSEP_HW=256 DSEP_REPS=10
cat $L
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c8.json; cut -c1-300 gpurun_out/bench_c8.json

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c15.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for d in 0 2; do run DSEP_RES=1 DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
for d in 0 2; do run DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
run DSEP_REPS=400
run DSEP_RES=1 DSEP_REPS=400
cat $L
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c15.json; cut -c1-300 gpurun_out/bench_c15.json

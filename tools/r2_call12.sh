#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c12.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for d in 0 1 2; do run DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
run DSEP_REPS=400
run DSEP_RES=1 DSEP_REPS=20
run DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CIN=256 DSEP_COUT=128 DSEP_HW=256 DSEP_REPS=10
cat $L
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c12.json; cut -c1-300 gpurun_out/bench_c12.json
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c12.csv python tools/profile_eval.py | tail -1

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c30.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=20
run DSEP_REPS=400
run DSEP_CIN=256 DSEP_REPS=20
cat $L
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_c30.log; tail -4 gpurun_out/pytest_gpu_c30.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c30.json; cut -c1-200 gpurun_out/bench_c30.json

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c19.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=20
run DSEP_REPS=400
run DSEP_RES=1 DSEP_REPS=400
cat $L
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu_c19.log; tail -6 gpurun_out/pytest_gpu_c19.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_c19.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c19.json; cut -c1-300 gpurun_out/bench_c19.json

#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "wide or fused8 or residual" 2>&1 | tail -3
L=gpurun_out/conv_modes_c31.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=20
run DSEP_RES=1 DSEP_REPS=400
cat $L
DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c31 python tools/profile_conv.py > /dev/null 2>&1
DSEP_RES=1 DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c31_res python tools/profile_conv.py > /dev/null 2>&1
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c31.csv python tools/profile_eval.py | tail -1
ls -la gpurun_out/*c31*

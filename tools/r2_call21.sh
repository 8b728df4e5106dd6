#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 > gpurun_out/pytest_gpu_c21.log; tail -6 gpurun_out/pytest_gpu_c21.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke_c21.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c21.json; cut -c1-300 gpurun_out/bench_c21.json
DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c21 python tools/profile_conv.py > /dev/null 2>&1
DSEP_RES=1 DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c21_res python tools/profile_conv.py > /dev/null 2>&1
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c21.csv python tools/profile_eval.py | tail -1
ls -la gpurun_out/*c21*

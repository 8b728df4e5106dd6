#!/usr/bin/env bash
# final validation of the shipped binary: tests, smoke, bench, launch list, memory-bound kernels, level breakdown
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_c58.log; tail -3 gpurun_out/pytest_gpu_c58.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee gpurun_out/smoke_c58.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c58.json; cut -c1-200 gpurun_out/bench_c58.json
DSEP_CUDA_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c58.csv python tools/profile_eval.py | tail -1
bash tools/ncu_membound.sh
timeout 300 python tools/profile_levels.py 8:265 8:18 18:33 33:48 48:63 63:86 86:101 101:143 143:163 163:186 186:205 205:225 225:245 245:265 48:205 > gpurun_out/levels_c58.log 2>&1; tail -3 gpurun_out/levels_c58.log

#!/usr/bin/env bash
# round-2 first GPU call: 2-unit (e4m3 correction) mode as the default — parity tests, bench, conv timing modes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu_c1.log; tail -5 gpurun_out/pytest_gpu_c1.log
python __graft_entry__.py --smoke 2>&1 | tail -1
for d in 0 1 2 4 6; do DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1; done > gpurun_out/conv_modes_p2_c1.log
for d in 0 1 2; do DSEP_PASSES=3 DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1; done >> gpurun_out/conv_modes_p2_c1.log
cat gpurun_out/conv_modes_p2_c1.log
python bench.py 2>&1 | tail -1 > gpurun_out/bench_c1.json; cut -c1-400 gpurun_out/bench_c1.json
bash tools/ncu_conv_modes.sh
python tools/parity_trajectory.py 30 2>&1 | tail -4 > gpurun_out/parity_traj_c1.log; cat gpurun_out/parity_traj_c1.log

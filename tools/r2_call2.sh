#!/usr/bin/env bash
# round-2 call 2: role-split fused conv kernel (conv_fused.cu) vs the previous halo kernel; graded-size parity tests
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu_c2.log; tail -4 gpurun_out/pytest_gpu_c2.log
L=gpurun_out/conv_modes_c2.log; : > $L
for d in 0 1 2 4; do echo "v2 debug=$d" >> $L; DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> $L; done
echo "v2 passes=3" >> $L; DSEP_PASSES=3 DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> $L
echo "v2 res" >> $L; DSEP_RES=1 DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> $L
echo "v2 sustained" >> $L; DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=400 python tools/profile_conv.py 2>&1 | tail -1 >> $L
echo "v1" >> $L; DSEP_CONV_V2=0 DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> $L
for d in 0 1 2; do echo "v1 2cta debug=$d" >> $L; DSEP_CONV_2CTA=1 DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> $L; done
cat $L
DSEP_CONV_2CTA=1 python -m pytest tests/test_ops_gpu.py tests/test_graded_gpu.py -q -m gpu -k "fused8 or e4m3 or level0" 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/bench_c2.json; cut -c1-300 gpurun_out/bench_c2.json

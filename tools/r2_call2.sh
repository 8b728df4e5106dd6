#!/usr/bin/env bash
# round-2 call 2: graded-size parity tests, cta_group::2 variant of the 2-unit conv
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_c2.log; tail -4 gpurun_out/pytest_gpu_c2.log
DSEP_CONV_2CTA=1 python -m pytest tests/test_ops_gpu.py tests/test_graded_gpu.py -q -m gpu -k "fused8 or e4m3 or level0" 2>&1 | tail -5
for d in 0 1 2; do DSEP_CONV_2CTA=1 DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1; done > gpurun_out/conv_modes_two_c2.log
DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=20 python tools/profile_conv.py 2>&1 | tail -1 >> gpurun_out/conv_modes_two_c2.log
DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=400 python tools/profile_conv.py 2>&1 | tail -1 >> gpurun_out/conv_modes_two_c2.log
DSEP_CONV_2CTA=1 DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_REPS=400 python tools/profile_conv.py 2>&1 | tail -1 >> gpurun_out/conv_modes_two_c2.log
cat gpurun_out/conv_modes_two_c2.log

#!/usr/bin/env bash
# A/B of the next-tile L2 prefetch in conv_wide (DSEP_CONV_DEBUG=8 switches it off)
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "tap_gather or narrow_conv or wide or fused8 or fused_prologue" 2>&1 | tail -3
for dbg in 8 0 8 0; do
echo "DSEP_CONV_DEBUG=$dbg"
for mode in "DSEP_REPS=20" "DSEP_RES=1 DSEP_REPS=20" "DSEP_SHORT=256 DSEP_REPS=20" "DSEP_SHORT=256 DSEP_REPS=300" "DSEP_REPS=300" "DSEP_K=1 DSEP_COUT=384 DSEP_REPS=20"; do
env $mode DSEP_CONV_DEBUG=$dbg DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 120 python tools/profile_conv.py 2>&1 | tail -1
done
done
for i in 1 2; do
DSEP_CONV_DEBUG=8 timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
done
} > gpurun_out/call54.log 2>&1

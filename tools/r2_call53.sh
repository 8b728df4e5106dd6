#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "tap_gather or narrow_conv" 2>&1 | tail -5
for i in 1 2; do
DSEP_PYR_TAPS_MIN=0 DSEP_FIR_KT=1 timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
done
} > gpurun_out/call53.log 2>&1

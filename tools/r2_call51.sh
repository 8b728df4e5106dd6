#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests -q -m gpu -x -k "fir or resample or upfirdn" 2>&1 | tail -3
for kt in 1 2 4 8; do echo KT=$kt; DSEP_FIR_KT=$kt timeout 120 python tools/profile_fir.py 2>&1 | tail -2; done
} > gpurun_out/call51.log 2>&1

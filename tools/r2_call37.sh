#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
DSEP_SHORT=256 DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c37_short python tools/profile_conv.py > /dev/null 2>&1
ls -la gpurun_out/conv_c37*

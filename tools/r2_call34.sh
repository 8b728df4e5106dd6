#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
for i in 1 2; do
python tools/profile_fir.py 2>&1 | tail -2
DSEP_LIB=diffsep_b200/build/variants/libdsep_fir3.so python tools/profile_fir.py 2>&1 | tail -2
done

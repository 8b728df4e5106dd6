#!/usr/bin/env bash
# The round-end check, as one command for `gpurun -- bash tools/gpu_check.sh`:
# GPU parity tests, smoke, the bench line, the reference arm, and the two ncu artefacts kept in profiles/.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py 2>&1 | tail -1 > gpurun_out/bench.json; cut -c1-300 gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv python tools/profile_eval.py | tail -1
DSEP_FUSEDIN=1 DSEP_STATS=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -f \
    -o gpurun_out/conv python tools/profile_conv.py | tail -1
# experimental e4m3-correction mode: build the variant HERE first
#   tools/build_variant.sh fp8corr -DDSEP_FP8_CORR=1
# then on the box:
#   V=diffsep_b200/build/variants/libdsep_fp8corr.so
#   DSEP_LIB=$V python -m pytest tests -q -m gpu -k "fused8_e4m3 or e4m3_correction_mode"
#   DSEP_LIB=$V DSEP_PASSES=2 DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py
#   DSEP_LIB=$V DSEP_PASSES=2 python bench.py

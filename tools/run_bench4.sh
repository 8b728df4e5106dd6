DSEP_DEBUG_SYNC=1 DSEP_BENCH_BATCH=1 python tools/profile_eval.py 2>&1 | tail -3
set -x
python tools/profile_conv.py
DSEP_RES=1 python tools/profile_conv.py
DSEP_PASSES=1 python tools/profile_conv.py
DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 python tools/profile_conv.py
python tools/profile_eval.py
DSEP_PASSES=1 python tools/profile_eval.py
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python tools/profile_eval.py | tail -2

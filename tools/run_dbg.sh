DSEP_DEBUG_SYNC=1 DSEP_BENCH_BATCH=1 python tools/profile_eval.py 2>&1 | tail -3

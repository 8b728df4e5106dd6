set -x
python __graft_entry__.py --smoke 2>&1 | tail -3
DSEP_BENCH_BATCH=4 python bench.py --steps 1 --warmup 1 2>&1 | tail -5

python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q -x 2>&1 | grep -E "^E  .*(Error|assert)|passed|failed|^FAILED" | cut -c1-300 | head -8
DSEP_DEBUG_SYNC=1 DSEP_BENCH_BATCH=4 python tools/profile_eval.py 2>&1 | grep -E "Cout_pad|one evaluation" | cut -c1-900

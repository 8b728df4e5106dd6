set -x
python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -5
python tools/profile_conv.py
DSEP_RES=1 python tools/profile_conv.py
DSEP_PASSES=1 python tools/profile_conv.py
python tools/profile_eval.py
DSEP_PASSES=1 python tools/profile_eval.py

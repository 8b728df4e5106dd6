python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -8
echo fuse=1; python tools/profile_eval.py
echo fuse=0; DSEP_FUSE=0 python tools/profile_eval.py
echo fuse=1 p1; DSEP_PASSES=1 python tools/profile_eval.py

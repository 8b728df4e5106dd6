python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -3
echo halo=1; python tools/profile_eval.py
echo halo=0; DSEP_CONV_HALO=0 python tools/profile_eval.py
echo halo=1 p1; DSEP_PASSES=1 python tools/profile_eval.py
echo halo=0 p1; DSEP_PASSES=1 DSEP_CONV_HALO=0 python tools/profile_eval.py

python -m pytest tests/test_ops_gpu.py -q -k "conv2d" 2>&1 | tail -2
DSEP_FUSEDIN=1 python tools/profile_conv.py
DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_RES=1 python tools/profile_conv.py
DSEP_FUSEDIN=1 DSEP_CONV_DEBUG=1 python tools/profile_conv.py
DSEP_FUSEDIN=1 DSEP_CIN=256 python tools/profile_conv.py
python tools/profile_eval.py

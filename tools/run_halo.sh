echo "== base_off unset"; python -m pytest tests/test_ops_gpu.py -q -k "conv2d" 2>&1 | tail -4
echo "== base_off set"; DSEP_CONV_DEBUG=8 python -m pytest tests/test_ops_gpu.py -q -k "conv2d" 2>&1 | tail -4
python tools/profile_conv.py
DSEP_CONV_DEBUG=8 python tools/profile_conv.py
DSEP_CONV_HALO=0 python tools/profile_conv.py

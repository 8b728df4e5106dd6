#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "priormix_with_network" > gpurun_out/c66_a.log 2>&1
DSEP_STFT_TC=0 timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "priormix_with_network" > gpurun_out/c66_b.log 2>&1
DSEP_LIB=diffsep_b200/build/variants/libdsep_prev.so timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -x -k "priormix_with_network" > gpurun_out/c66_c.log 2>&1
tail -3 gpurun_out/c66_a.log gpurun_out/c66_b.log gpurun_out/c66_c.log

#!/usr/bin/env bash
mkdir -p gpurun_out
{ T=8000 timeout 300 python tools/debug/dft_debug.py; T=4096 timeout 300 python tools/debug/dft_debug.py; DSEP_CUDA_GRAPH=0 T=8000 timeout 300 python tools/debug/dft_debug.py; } > gpurun_out/c67.log 2>&1

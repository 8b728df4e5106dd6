#!/usr/bin/env bash
mkdir -p gpurun_out
{ T=8000 timeout 300 python tools/debug/dft_debug2.py; } > gpurun_out/c68.log 2>&1

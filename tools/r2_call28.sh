#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/parity_report.py > gpurun_out/parity_c28.md 2> gpurun_out/parity_c28.err; tail -3 gpurun_out/parity_c28.err
timeout 1200 python tools/parity_trajectory.py 30 >> gpurun_out/parity_c28.md 2>> gpurun_out/parity_c28.err; tail -3 gpurun_out/parity_c28.err
cat gpurun_out/parity_c28.md

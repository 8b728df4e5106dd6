#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
L=gpurun_out/sanitizer_c39.log; : > $L
echo "== memcheck: conv_wide / conv_fused (2-unit), im2col, fp32 FIR, attention_tc, training-side kernels" >> $L
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "wide_tiles or fused8 or im2col or input_conv or fir_resample_f32 or training_forward or (attention and 2x256x256)" 2>&1 | tail -8 >> $L
echo "exit=$?" >> $L
echo "== racecheck: conv_wide (shared-memory hazards between builders / TMA / epilogue)" >> $L
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "wide_tiles and shape0" 2>&1 | tail -8 >> $L
echo "exit=$?" >> $L
cat $L

#!/usr/bin/env bash
# compute-sanitizer on the last session's kernels: fir_tile_kernel (several tiles per block), tap_gather_kernel,
# the 1x1 + gather form of the narrow conv
mkdir -p gpurun_out
L=gpurun_out/sanitizer_c59.log; : > $L
echo "== memcheck: fir_tile (tiles-per-block loop), tap_gather, narrow conv as 1x1 + gather" >> $L
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "fir or tap_gather or narrow_conv" 2>&1 | tail -8 >> $L
echo "exit=$?" >> $L
echo "== racecheck: fir_tile (shared-memory patch re-staged per tile)" >> $L
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "fir_fused_groupnorm_branch or fir_resample_f32" 2>&1 | tail -8 >> $L
echo "exit=$?" >> $L
cat $L

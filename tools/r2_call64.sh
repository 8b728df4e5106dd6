#!/usr/bin/env bash
# attention_tc_kernel with software-pipelined operand loads vs the previous build (DSEP_LIB=.../libdsep_prev.so)
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "attention" 2>&1 | tail -3
for lib in diffsep_b200/build/variants/libdsep_prev.so ""; do
echo "DSEP_LIB=$lib"
for shape in "32 256 256" "8 1920 256" "16 496 256" "32 256 128"; do
DSEP_LIB=$lib timeout 120 python tools/profile_attention.py $shape 2>&1 | tail -1
done
done
} > gpurun_out/call64.log 2>&1

#!/usr/bin/env bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for pdl in 0 1; do
echo "DSEP_PDL=$pdl"
DSEP_PDL=$pdl timeout 600 python tools/profile_levels.py 8:265 48:205 63:186 245:265
done
for i in 1 2; do
DSEP_PDL=0 timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
timeout 600 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | cut -c1-130
done
} > gpurun_out/call56.log 2>&1

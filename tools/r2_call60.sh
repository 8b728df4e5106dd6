#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
DSEP_BENCH_WORKLOAD="configs[4]" timeout 900 python bench.py --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c60_cfg4.json; cut -c1-200 gpurun_out/bench_c60_cfg4.json
DSEP_BENCH_WORKLOAD="configs[3]" timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c60_cfg3.json; cut -c1-200 gpurun_out/bench_c60_cfg3.json

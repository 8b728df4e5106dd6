#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu_c10.log; tail -6 gpurun_out/pytest_gpu_c10.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_c10.json; cut -c1-300 gpurun_out/bench_c10.json
DSEP_BENCH_WORKLOAD="configs[4]" timeout 900 python bench.py --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c10_cfg4.json; cut -c1-300 gpurun_out/bench_c10_cfg4.json
DSEP_BENCH_WORKLOAD="configs[3]" timeout 900 python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c10_cfg3.json; cut -c1-300 gpurun_out/bench_c10_cfg3.json

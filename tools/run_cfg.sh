DSEP_BENCH_WORKLOAD="configs[3]" python bench.py --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_cfg3_r1.json; cut -c1-200 gpurun_out/bench_cfg3_r1.json
DSEP_BENCH_WORKLOAD="configs[4]" python bench.py --steps 1 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_cfg4_r1.json; cut -c1-200 gpurun_out/bench_cfg4_r1.json

set -x
python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q 2>&1 | tail -5
python tools/profile_conv.py
python tools/profile_eval.py
python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1c.json
cat gpurun_out/bench_r1c.json
DSEP_PASSES=1 python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1c_p1.json
cat gpurun_out/bench_r1c_p1.json

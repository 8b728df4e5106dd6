timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "groupnorm or fir or combine or attention or sde or normalize or sgemm or time_embedding" 2>&1 | tail -6
echo "exit=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -q -x -k "conv2d_tc and 16x24" 2>&1 | tail -6

set -x
python tools/profile_eval.py
python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_b32_first.json
cat gpurun_out/bench_b32_first.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1a.csv python tools/profile_eval.py | tail -2

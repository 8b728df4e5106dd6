set -x
python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1d.json
cut -c1-300 gpurun_out/bench_r1d.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_eval.py | tail -1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -f -o gpurun_out/conv_r1d python tools/profile_conv.py | tail -1

set -x
python bench.py --steps 3 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1f.json
cut -c1-250 gpurun_out/bench_r1f.json
DSEP_FUSEDIN=1 DSEP_STATS=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -f -o gpurun_out/conv_r1f python tools/profile_conv.py | tail -1
DSEP_PASSES=1 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r1f_p1.json
cut -c1-250 gpurun_out/bench_r1f_p1.json

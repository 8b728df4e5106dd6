#!/usr/bin/env bash
# round-2 call 4: CTA-pair MMA variant of the role-split fused conv
set -x
mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_graded_gpu.py -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_gpu_c4.log; tail -4 gpurun_out/pytest_gpu_c4.log
L=gpurun_out/conv_modes_c4.log; : > $L
run() { echo "$*" >> $L; env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
for d in 0 1 2 4; do run DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
run DSEP_REPS=400
run DSEP_RES=1 DSEP_REPS=20
for d in 0 1 2; do run DSEP_CONV_PAIR=0 DSEP_CONV_DEBUG=$d DSEP_REPS=20; done
run DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CONV_PAIR=0 DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CONV_V2=0 DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
cat $L
python -m pytest tests -q -m gpu 2>&1 | tail -5
python bench.py 2>&1 | tail -1 > gpurun_out/bench_c4.json; cut -c1-300 gpurun_out/bench_c4.json
M=gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum
DSEP_FUSEDIN=1 DSEP_STATS=1 ncu --metrics $M --clock-control none -k regex:conv_ -s 3 -c 1 --csv --log-file gpurun_out/conv_c4_metrics.csv python tools/profile_conv.py > /dev/null 2>&1
DSEP_FUSEDIN=1 DSEP_STATS=1 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c4 python tools/profile_conv.py > /dev/null 2>&1
ls -la gpurun_out/conv_c4*

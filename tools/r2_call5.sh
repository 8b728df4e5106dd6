#!/usr/bin/env bash
# round-2 call 5: ncu of the role-split conv (1-CTA MMAs), launch list of one evaluation, pair-mode diagnostics under timeouts
set -x
mkdir -p gpurun_out
L=gpurun_out/conv_modes_c5.log; : > $L
run() { echo "$*" >> $L; timeout 120 env "$@" DSEP_FUSEDIN=1 DSEP_STATS=1 python tools/profile_conv.py 2>&1 | tail -1 >> $L; }
run DSEP_REPS=20
run DSEP_CONV_PAIR=1 DSEP_REPS=20
run DSEP_CONV_PAIR=1 DSEP_CONV_DEBUG=2 DSEP_REPS=20
run DSEP_CONV_PAIR=1 DSEP_CONV_DEBUG=1 DSEP_REPS=20
run DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CONV_V2=0 DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 DSEP_REPS=20
run DSEP_CIN=256 DSEP_COUT=128 DSEP_HW=256 DSEP_REPS=10
run DSEP_CONV_V2=0 DSEP_CIN=256 DSEP_COUT=128 DSEP_HW=256 DSEP_REPS=10
cat $L
M=gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum
for d in 0 1 2; do
DSEP_CONV_DEBUG=$d DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 300 ncu --metrics $M --clock-control none -k regex:conv_ -s 3 -c 1 --csv --log-file gpurun_out/conv_c5_metrics_$d.csv python tools/profile_conv.py > /dev/null 2>&1
done
DSEP_FUSEDIN=1 DSEP_STATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 3 -c 1 -f -o gpurun_out/conv_c5 python tools/profile_conv.py > /dev/null 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5.csv python tools/profile_eval.py | tail -1
ls -la gpurun_out/ | tail -8

#!/usr/bin/env bash
# ncu counters of the level-0 fused conv in its three timing modes (full / no MMA issue / no patch builds)
# -> gpurun_out/conv_modes_<d>.csv.  Usage: gpurun -- bash tools/ncu_conv_modes.sh [DSEP_LIB path]
M=gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,lts__t_bytes.sum.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum
mkdir -p gpurun_out
for d in 0 1 2; do
  DSEP_LIB=$1 DSEP_FUSEDIN=1 DSEP_STATS=1 DSEP_CONV_DEBUG=$d ncu --metrics $M --clock-control none -k regex:conv_tc -s 3 -c 1 --csv --log-file gpurun_out/conv_modes_$d.csv python tools/profile_conv.py > gpurun_out/conv_modes_$d.log 2>&1
  tail -2 gpurun_out/conv_modes_$d.log
done

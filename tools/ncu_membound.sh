#!/usr/bin/env bash
# achieved HBM traffic / throughput of the memory-bound kernels of one PC step -> gpurun_out/membound.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none \
  -k regex:'fir_|tap_gather|im2col|spec_pack|stft_frames|istft_ola|gn_act_split|combine|channel_stats|sde_update|sgemm|out_head|attention|normalize|scale_output|gn_tables|time_embedding|film' \
  --csv --log-file gpurun_out/membound.csv python tools/profile_step.py | tail -1

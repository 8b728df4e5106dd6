ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_eval.py | tail -1

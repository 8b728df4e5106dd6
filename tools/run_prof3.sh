set -x
python tools/profile_conv.py
DSEP_RES=1 python tools/profile_conv.py
DSEP_PASSES=1 python tools/profile_conv.py
DSEP_CIN=256 DSEP_COUT=256 DSEP_HW=64 python tools/profile_conv.py
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -f -o gpurun_out/conv_r1a python tools/profile_conv.py | tail -2
DSEP_RES=1 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -f -o gpurun_out/conv_res_r1a python tools/profile_conv.py | tail -2

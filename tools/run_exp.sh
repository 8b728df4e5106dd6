for d in 0 1 2 3 4 5 6; do echo "debug=$d"; DSEP_CONV_DEBUG=$d python tools/profile_conv.py; done
for d in 0 1 2 4; do echo "p1 debug=$d"; DSEP_PASSES=1 DSEP_CONV_DEBUG=$d python tools/profile_conv.py; done

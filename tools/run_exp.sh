for d in 0 1 2 3; do echo "halo debug=$d"; DSEP_CONV_DEBUG=$d python tools/profile_conv.py; done
for d in 0 1 2; do echo "halo p1 debug=$d"; DSEP_PASSES=1 DSEP_CONV_DEBUG=$d python tools/profile_conv.py; done

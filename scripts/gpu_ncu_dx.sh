#!/bin/bash
# ncu --set full (with SASS-level stall sampling) of the three conv3x3_dx_kernel launches of one forward at the bench shape.
set -u
mkdir -p gpurun_out
CMD="python scripts/ab_conv_modes.py 64"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_dx -s 6 -c 3 -o gpurun_out/prof_dx -f $CMD > gpurun_out/ncu_dx.log 2>&1
echo "dx full exit $?"
ls -la gpurun_out/prof_dx.ncu-rep

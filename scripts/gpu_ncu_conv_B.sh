#!/bin/bash
# ncu --set full of the 17 conv3x3_halo_kernel / conv3x3_dx_kernel launches of ONE forward at the bench's default workload (variant B, batch 64)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0"
$CMD > gpurun_out/ncu_plain_B.json 2> gpurun_out/ncu_plain_B.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_B.err; exit 1; }
timeout 900 ncu --set full --clock-control none -k "regex:conv3x3_(halo|dx)" -s 34 -c 17 -o gpurun_out/prof_conv_B -f $CMD > gpurun_out/ncu_conv_B.log 2>&1
echo "conv full exit $?"
ls -la gpurun_out/prof_conv_B.ncu-rep

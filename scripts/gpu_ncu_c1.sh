#!/bin/bash
# ncu --set full of the first-layer conv (conv3x3_c1_kernel) at the bench shape (variant B).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0 --c2-clips 0"
timeout 300 $CMD > gpurun_out/ncu_c1_plain.json 2> gpurun_out/ncu_c1_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_c1_plain.err; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3x3_c1 -s 4 -c 1 -o gpurun_out/prof_c1 -f $CMD > gpurun_out/ncu_c1.log 2>&1
echo "c1 full exit $?"
ls -la gpurun_out/prof_c1.ncu-rep

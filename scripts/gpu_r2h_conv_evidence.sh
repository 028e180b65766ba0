#!/bin/bash
# ncu --set full of the 17 tensor-core conv launches of ONE forward of the FINAL build (halo / dx / upm / upm2 incl. the fused-pool variant)
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 --c2-clips 0"
timeout 300 $CMD > gpurun_out/ncu_plain_B.json 2> gpurun_out/ncu_plain_B.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_B.err; exit 1; }
timeout 900 ncu --set full --clock-control none -k "regex:conv3x3_(halo|dx|upm)" -s 34 -c 17 -o gpurun_out/prof_conv_B -f $CMD > gpurun_out/ncu_conv_B.log 2>&1
echo "conv full exit $?"
python scripts/ncu_summary.py gpurun_out/prof_conv_B.ncu-rep gpurun_out/r2h_conv_full_variantB.csv > /dev/null
rm -f gpurun_out/prof_conv_B.ncu-rep
cut -c1-200 gpurun_out/r2h_conv_full_variantB.csv

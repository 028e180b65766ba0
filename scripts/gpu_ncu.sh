#!/bin/bash
# ncu evidence for bench.py (variant R keeps replays short): launch list + full sections of the conv kernel; full sections of the
# weight-gradient GEMM from the training driver.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --variant R --no-cpu-baseline --train-steps 0"
$CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:conv3x3_(halo|dx)" -s 48 -c 17 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
echo "conv full exit $?"
python scripts/prof_train.py > gpurun_out/prof_train_plain.log 2>&1 || { echo "train plain failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 50 -c 25 -o gpurun_out/prof_wgrad -f python scripts/prof_train.py > gpurun_out/ncu_wgrad.log 2>&1
echo "wgrad full exit $?"
ls -la gpurun_out/*.ncu-rep

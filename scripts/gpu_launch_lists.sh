#!/bin/bash
# ncu launch lists (gpu__time_duration.sum, cold-cache, serialised) of the inference bench (variant R keeps replays short) and of
# the training-step driver, each after a plain run of the same command has exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --variant R --no-cpu-baseline --train-steps 0 --c2-clips 0"
timeout 300 $CMD > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "inference launch list exit $?"
timeout 300 python scripts/prof_train.py > gpurun_out/prof_train_plain.log 2>&1 || { echo "train plain failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python scripts/prof_train.py > gpurun_out/ncu_launches_train.log 2>&1
echo "train launch list exit $?"
wc -l gpurun_out/launches.csv gpurun_out/launches_train.csv

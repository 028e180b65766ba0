#!/bin/bash
# training step: tests, then the bench's train_step object with the weight gradients on two / one / no side streams
set -u
mkdir -p gpurun_out
for ov in 1 0 1 0; do
ADN_SPLIT_PACK=$ov timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c2-clips 0 > gpurun_out/bench_train_$ov.json 2> gpurun_out/bench_train.err; echo "split pack=$ov bench exit $?"; tail -3 gpurun_out/bench_train.err
python -c "
import json; d=json.load(open('gpurun_out/bench_train_$ov.json')); t=d['train_step']; print({k:t[k] for k in ('ms_per_step','pairs_per_s','kernel_launches_per_step','tflops')}, t['batch64'])"
done

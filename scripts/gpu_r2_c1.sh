#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -m gpu -x -q -k "first_layer or fixture or checkpoints or c1 or train_step_matches or forward_and_backward" > gpurun_out/pytest_c1.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_c1.log
ADN_C1_SETS=2 timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -x -q -k "first_layer or fixture" 2>&1 | tail -1
for rep in 1 2; do
for impl in tc1 tc2 fma; do
case $impl in tc1) export ADN_C1_IMPL=tc ADN_C1_SETS=1;; tc2) export ADN_C1_IMPL=tc ADN_C1_SETS=2;; fma) export ADN_C1_IMPL=fma;; esac
timeout 600 python bench.py --steps 5 --warmup 3 --layers --no-cpu-baseline --c2-clips 0 --train-steps 0 > gpurun_out/bench_c1_$impl.json 2> gpurun_out/bench_c1_$impl.err; 
python -c "
import json; d=json.load(open('gpurun_out/bench_c1_$impl.json')); print('$impl', round(d['ms_per_step'],3), round(d['kernels']['conv3x3_c1']['ms_per_step'],4), round(d['kernels']['layers']['downconv1.3']['ms'],3))"
done; done

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py tests/test_gpu_train_step.py -m gpu -q -s > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; grep -E "^(linear|real) loss|passed|failed|Error" gpurun_out/pytest_train.log | head -20
bash scripts/gpu_bench.sh
python scripts/prof_train.py > gpurun_out/prof_train_plain.log 2>&1 || { echo plain failed; tail -5 gpurun_out/prof_train_plain.log; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python scripts/prof_train.py > gpurun_out/ncu_train.log 2>&1; echo "ncu exit $?"

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train_step.py -m gpu -q -s > gpurun_out/pytest_train_step.log 2>&1; echo "pytest exit $?"; grep -E "^shape|passed|failed|Error|error" gpurun_out/pytest_train_step.log | head -30; tail -30 gpurun_out/pytest_train_step.log

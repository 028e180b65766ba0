#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/wgrad_sweep.py > gpurun_out/wgrad_sweep.log 2>&1; echo "sweep exit $?"; cat gpurun_out/wgrad_sweep.log | tail -14
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q > gpurun_out/pytest_train_ops.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_train_ops.log

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_ops.py -m gpu -q > gpurun_out/pytest_train_ops.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_train_ops.log

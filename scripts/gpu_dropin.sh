#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_dropin.py -m gpu -q -x > gpurun_out/pytest_dropin.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_dropin.log

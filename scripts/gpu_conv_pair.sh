#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_unet.py -m gpu -q -x > gpurun_out/pytest_unet.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_unet.log

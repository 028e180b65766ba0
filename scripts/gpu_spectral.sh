#!/bin/bash
# spectral kernels only: parity tests + microbench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/pytest_spectral.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_spectral.log
timeout 300 python scripts/bench_spectral.py 20 > gpurun_out/spectral.log 2>&1; echo "spectral exit $?"; cat gpurun_out/spectral.log

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py -m gpu -x -q > gpurun_out/pytest_spectral.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_spectral.log
timeout 300 python scripts/bench_spectral.py 20 > gpurun_out/spectral.log 2>&1; echo "spectral exit $?"; cat gpurun_out/spectral.log
ncu --set full --clock-control none --import-source on -k regex:"${1:-stft_kernel}" -s 2 -c 1 -o gpurun_out/prof_${2:-stft} -f python scripts/prof_spectral.py > gpurun_out/ncu_${2:-stft}.log 2>&1; echo "ncu exit $?"

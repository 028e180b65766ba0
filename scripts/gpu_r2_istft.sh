#!/bin/bash
# round 2: warp-specialised iSTFT -- parity tests, then the config-2 microbench with both implementations
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/pytest_spectral.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_spectral.log
ADN_ISTFT_IMPL=2 timeout 300 python scripts/bench_spectral.py 20 > gpurun_out/spectral_ws.log 2>&1; echo "spectral (ws) exit $?"; grep -E "stft" gpurun_out/spectral_ws.log
ADN_ISTFT_IMPL=1 timeout 300 python scripts/bench_spectral.py 20 > gpurun_out/spectral_v1.log 2>&1; echo "spectral (v1) exit $?"; grep -E "istft" gpurun_out/spectral_v1.log

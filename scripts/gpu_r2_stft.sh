#!/bin/bash
# round 2: spectral tests on both STFT implementations + config-2 timings
set -u
mkdir -p gpurun_out
for impl in 2 1; do
  ADN_STFT_IMPL=$impl timeout 600 python -m pytest tests/test_gpu_spectral.py tests/test_gpu_data_loader.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/pytest_spectral_impl$impl.log 2>&1; echo "impl $impl pytest exit $?"; tail -4 gpurun_out/pytest_spectral_impl$impl.log
  ADN_STFT_IMPL=$impl timeout 300 python scripts/bench_spectral.py 20 2>&1 | grep -E "^stft" | sed "s/^/impl=$impl /"
done

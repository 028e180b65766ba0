#!/bin/bash
# quick iteration: selected tests + spectral microbench + short bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_spectral.py 20 > gpurun_out/spectral.log 2>&1; echo "spectral exit $?"; cat gpurun_out/spectral.log
timeout 600 python bench.py --steps 5 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench B exit $?"; tail -3 gpurun_out/bench_B.err

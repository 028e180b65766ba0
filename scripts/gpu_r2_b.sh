#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR|vs reference" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python scripts/bench_spectral.py 20 2>&1 | grep -E "^stft|^istft"
timeout 900 python bench.py --steps 5 --warmup 3 --layers --train-steps 0 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python scripts/show_bench.py gpurun_out/bench_n1.json 2>/dev/null | head -60

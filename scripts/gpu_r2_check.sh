#!/bin/bash
# round 2: full GPU suite + STFT CTAs-per-SM experiment + short bench (parity / config-2 leg)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu.log
for c in 1 2; do
  ADN_STFT_CTAS_PER_SM=$c timeout 300 python scripts/bench_spectral.py 20 2>&1 | grep -E "^stft_c2|^istft" | sed "s/^/ctas=$c /"
done
timeout 900 python bench.py --steps 5 --warmup 3 --layers > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python scripts/show_bench.py gpurun_out/bench_n1.json 2>/dev/null | head -60

#!/bin/bash
# round 2: what the driver runs at round end (GPU tests, smoke, default bench, reference arm) + launch list
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --layers > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -3 gpurun_out/bench_default.err
python scripts/show_bench.py gpurun_out/bench_default.json 2>/dev/null | head -70
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; cut -c1-600 gpurun_out/bench_reference.json

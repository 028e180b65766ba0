#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both clip variants), ncu launch list.  Logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench B exit $?"; tail -c 6000 gpurun_out/bench_B.json; tail -5 gpurun_out/bench_B.err
timeout 600 python bench.py --steps 5 --warmup 3 --layers --variant R --no-cpu-baseline > gpurun_out/bench_R.json 2> gpurun_out/bench_R.err; echo "bench R exit $?"; tail -c 6000 gpurun_out/bench_R.json; tail -5 gpurun_out/bench_R.err

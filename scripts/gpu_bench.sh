#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --layers > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench B exit $?"; tail -5 gpurun_out/bench_B.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_B.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','cpu_baseline')})
print(d['train_step'])
PY

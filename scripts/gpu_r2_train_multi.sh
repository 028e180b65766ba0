#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_train_step.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --train-steps 20 --c2-clips 0 > gpurun_out/bench_train_n$N.json 2> gpurun_out/bench_train_n$N.err; echo "bench N=$N exit $?"; tail -3 gpurun_out/bench_train_n$N.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_train_n$N.json')); t=d['train_step']; print('N=$N', {k:t[k] for k in ('ms_per_step','pairs_per_s','kernel_launches_per_step','parallelism')}, t['batch64'])"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --c2-clips 0 --train-steps 20 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "bench N=1 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); t=d['train_step']; print('N=1', {k:t[k] for k in ('ms_per_step','pairs_per_s','kernel_launches_per_step')}, t['batch64'])"

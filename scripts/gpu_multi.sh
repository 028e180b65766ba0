#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_unet.py -m gpu -x -q -k "fixture or conv3x3" > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_quick.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --layers > gpurun_out/bench_B_n$N.json 2> gpurun_out/bench_B_n$N.err; echo "bench N=$N exit $?"; tail -5 gpurun_out/bench_B_n$N.err | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 --layers --no-cpu-baseline > gpurun_out/bench_B.json 2> gpurun_out/bench_B.err; echo "bench N=1 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --variant R > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref.err; echo "ref arm exit $?"; cat gpurun_out/bench_ref_n$N.json | cut -c1-400

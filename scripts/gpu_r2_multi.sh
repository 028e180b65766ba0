#!/bin/bash
# round 2: N-GPU bench (copy-engine gather) + ADN_GATHER=nccl comparison
set -u
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
for mode in peer nccl; do
  ADN_GATHER=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --train-steps 0 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "bench N=$N $mode exit $?"; tail -4 gpurun_out/bench_n${N}_$mode.err | cut -c1-400
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}_$mode.json"))
print("$mode", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("communication"), d.get("gather_check"))
PY
done

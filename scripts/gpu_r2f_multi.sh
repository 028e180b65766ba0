#!/bin/bash
# the driver's N-GPU command (default gather, training leg included), then the line's key figures
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err; echo "bench N=$N exit $?"; tail -4 gpurun_out/bench_n${N}.err | cut -c1-400
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n${N}.json"))
t=d.get("train_step") or {}
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("communication"), d.get("gather_check"))
print("train", {k: t.get(k) for k in ("ms_per_step", "pairs_per_s")}, t.get("batch64"))
PY

#!/bin/bash
# 2+-GPU check of the sharded modes (bit-identical to 1 GPU) and bench at N GPUs
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/shard_check$N.log 2>&1; grep -v "^\*\|OMP\|^$" gpurun_out/shard_check$N.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["gpu_launches"])
PY

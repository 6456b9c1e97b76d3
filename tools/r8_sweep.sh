#!/bin/bash
# one 8-GPU box: correctness at 8 ranks (all three sharded modes), then fused vs replicate at N = 8, fused at N = 4
mkdir -p gpurun_out
run() { # N tag extra-args...
  N=$1; TAG=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((RANDOM % 90 + 10)) bench.py --gpus $N --steps 8 --warmup 3 "$@" > gpurun_out/sw_${TAG}.json 2> gpurun_out/sw_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_${TAG}.json").read().strip().splitlines()[-1]); print("${TAG}", round(d["value"]), "e2e", d["e2e"]["value"])
except Exception as e: print("${TAG}", "ERR", e)
PY
}
S3D_CHECK_FRAMES=70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/shard_check8.log 2>&1; grep "sharded_check\|Error" gpurun_out/shard_check8.log | tail -4
run 8 n8_fused
run 8 n8_replicate --shard-mode replicate --no-e2e
run 4 n4_fused --no-e2e
run 4 n4_replicate --shard-mode replicate --no-e2e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599 tools/trace_pipeline.py 2>&1 | grep -v "^\*\|OMP" | tail -3

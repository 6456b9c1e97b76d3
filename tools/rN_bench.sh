#!/bin/bash
# usage: tools/rN_bench.sh N  -- correctness of the sharded modes at N GPUs, then bench (fused with e2e, replicate without)
N=$1
mkdir -p gpurun_out
S3D_CHECK_FRAMES=${S3D_CHECK_FRAMES:-70} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/shard_check$N.log 2>&1; grep "sharded_check\|Error" gpurun_out/shard_check$N.log | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --shard-mode replicate --no-e2e > gpurun_out/bench_n${N}_rep.json 2> gpurun_out/bench_n${N}_rep.err || tail -5 gpurun_out/bench_n${N}_rep.err | cut -c1-300
python - <<PY
import json
for tag in ("", "_rep"):
    try:
        d=json.loads(open("gpurun_out/bench_n$N%s.json" % tag).read().strip().splitlines()[-1])
        print("N=$N", tag or "fused", round(d["value"]), "e2e", d["e2e"]["value"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"])
    except Exception as e: print("N=$N", tag, "ERR", e)
PY

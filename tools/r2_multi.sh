#!/bin/bash
# usage: tools/r2_multi.sh N [extra bench args]  -- N-GPU parity check of the sharded modes, then the sharded bench (chunk split and beam split)
N=$1; shift
mkdir -p gpurun_out
S3D_CHECK_FRAMES=${S3D_CHECK_FRAMES:-150} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/shard_check$N.log 2>&1; grep "sharded_check\|Error" gpurun_out/shard_check$N.log | tail -4
for split in chunks beams; do
  S3D_ROUTE_SPLIT=$split timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 6 --warmup 2 "$@" > gpurun_out/bench_n${N}_$split.json 2> gpurun_out/bench_n${N}_$split.err || { grep -v "^W1018\|^\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_$split.err | grep -i "error" | head -5; }
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$split.json").read().strip().splitlines()[-1])
    print("N=$N $split", round(d["value"]), "e2e", round(d.get("e2e",{}).get("value",0)), "single", round(d.get("single_gpu_same_workload",{}).get("value",0)), "speedup %.2f" % d.get("speedup_vs_one_gpu_same_workload",0), "nvlink GB/s/rank %.1f" % d["config"]["nvlink_GBps_per_rank"], "parity", d.get("parity",{}).get("ok"), d.get("parity",{}).get("sharded_vs_single_gpu"))
except Exception as e: print("N=$N $split ERR", e)
PY
done

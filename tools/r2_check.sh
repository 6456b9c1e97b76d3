#!/bin/bash
# 2-GPU check of both sharded modes + bench at N=2, and the e2e probe on one GPU
mkdir -p gpurun_out
python tools/e2e_probe.py > gpurun_out/e2e_probe.log 2>&1; tail -2 gpurun_out/e2e_probe.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/shard_check2.log 2>&1; grep -v "^\*\|OMP\|^$" gpurun_out/shard_check2.log | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2_rep.json 2> gpurun_out/bench_n2_rep.err; tail -c 1500 gpurun_out/bench_n2_rep.json; tail -3 gpurun_out/bench_n2_rep.err

#!/bin/bash
# usage: tools/prof1.sh <tag> <kernel-regex>   -- one warm-cache full capture of the named kernels on the bench workload
TAG=$1; K=$2
mkdir -p gpurun_out
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --frames-per-step 64"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$K" -s 12 -c 2 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_f_${TAG}.log

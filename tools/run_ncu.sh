#!/bin/bash
# usage: tools/run_ncu.sh <kernel-regex> <skip> <count> <tag> [bench args...]
set -e
K=$1; S=$2; C=$3; TAG=$4; shift 4
mkdir -p gpurun_out
CMD="python bench.py --no-e2e --no-cpu-baseline $*"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_l_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
tail -2 gpurun_out/plain_${TAG}.log

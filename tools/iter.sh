#!/bin/bash
# usage: tools/iter.sh <tag> [kernel-regex]  -- GPU parity tests, bench line, and (optionally) one warm-cache ncu capture
TAG=$1; K=$2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/b_${TAG}.json 2> gpurun_out/b_${TAG}.err || tail -c 600 gpurun_out/b_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_${TAG}.json"))
print("${TAG}", round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v["us_per_launch"],1) for k,v in d["roofline"]["kernels"].items()}, "retries", d["config"]["chunk_retries"], "frac %.4f" % d["roofline"]["frac"])
PY
[ -n "$K" ] && bash tools/prof1.sh ${TAG} "$K"

#!/bin/bash
# usage: tools/r2_variants.sh <lib> [<lib> ...]  -- cfg2 / cfg3 bench lines (with the bench's parity block) for other builds of the library (variants/*.so)
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename $lib .so)
  for wl in cfg2 cfg3; do
    S3D_LIB_PATH=$PWD/$lib timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-cfg3 --steps 10 --warmup 3 > gpurun_out/var_${tag}_$wl.json 2> gpurun_out/var_${tag}_$wl.err || tail -c 600 gpurun_out/var_${tag}_$wl.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_${tag}_$wl.json").read().strip().splitlines()[-1])
    print("$tag $wl", round(d["value"]), {k:round(v["us_per_launch_exclusive"],1) for k,v in d["roofline"]["kernels"].items()}, "parity", d.get("parity",{}).get("ok"))
except Exception as e: print("$tag $wl ERR", e)
PY
  done
done

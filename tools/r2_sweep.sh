#!/bin/bash
# usage: tools/r2_sweep.sh <tag> -- cfg2 bench under a few env settings (value only)
TAG=$1
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 8 --warmup 3 > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err || tail -c 300 gpurun_out/${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$name.json").read().strip().splitlines()[-1])
    print("${TAG} $name", round(d["value"]), {k:round(v["us_per_launch"],1) for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("$name ERR", e)
PY
}
if [ -n "$SWEEP" ]; then eval "$SWEEP"; exit 0; fi
run base A=1

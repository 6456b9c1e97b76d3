#!/bin/bash
# usage: tools/r2_apply_ab.sh <tag>  -- GPU parity tests, then cfg2 / cfg3 bench lines for 2 and 3 update-kernel blocks per SM, then the kernel's phase timers
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for bps in 3; do
  for wl in cfg2 cfg3; do
    S3D_APPLY_BPS=$bps timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-cfg3 --steps 10 --warmup 3 > gpurun_out/${TAG}_${wl}_b$bps.json 2> gpurun_out/${TAG}_${wl}_b$bps.err || tail -c 600 gpurun_out/${TAG}_${wl}_b$bps.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${wl}_b$bps.json").read().strip().splitlines()[-1])
    print("${TAG} bps=$bps $wl", round(d["value"]), {k:round(v["us_per_launch_exclusive"],1) for k,v in d["roofline"]["kernels"].items()}, "frac %.4f"%d["roofline"]["frac"], "parity", d.get("parity",{}).get("ok"), "unres", round(d.get("unreserved",{}).get("value",0)))
except Exception as e: print("$wl ERR", e)
PY
  done
done
# (the phase-timer build must exist: nvcc ... -DS3D_AP_PHASES -o variants/libsonar3d_phases.so, built in the container before the call)
[ -f variants/libsonar3d_phases.so ] && S3D_LIB_PATH=$PWD/variants/libsonar3d_phases.so timeout 250 python bench.py --workload cfg2 --no-cpu-baseline --no-e2e --no-cfg3 --steps 6 --warmup 3 2>&1 >/dev/null | grep phases | cut -c1-400

#!/bin/bash
# usage: tools/r2_iter.sh <tag> [pytest-args]  -- smoke, GPU parity tests, cfg2/cfg3 bench lines (each under its own timeout)
TAG=$1; shift
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 || { echo "SMOKE FAILED/HUNG"; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15
for wl in cfg2 cfg3; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-e2e --no-cfg3 > gpurun_out/${TAG}_$wl.json 2> gpurun_out/${TAG}_$wl.err || tail -c 800 gpurun_out/${TAG}_$wl.err
done
python - <<PY
import json
for f in ("cfg2","cfg3"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_{f}.json").read().strip().splitlines()[-1])
        print("${TAG}", f, round(d["value"]), {k:round(v["us_per_launch_exclusive"],1) for k,v in d["roofline"]["kernels"].items()}, "upd/s %.3g"%d["voxel_updates_per_s"], "frac %.4f"%d["roofline"]["frac"], "whole %.4f"%d["roofline"]["whole_path"]["frac"], "retries", d["config"].get("chunk_retries"), "parity", d.get("parity",{}).get("ok"))
    except Exception as e: print(f,"ERR",e)
PY

#!/bin/bash
# round-2 baseline: GPU tests + cfg2 / cfg3 bench lines of the round-1 code
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --steps 8 --warmup 3 > gpurun_out/r2_base_cfg2.json 2> gpurun_out/r2_base_cfg2.err || tail -c 600 gpurun_out/r2_base_cfg2.err
python bench.py --workload cfg3 --steps 4 --warmup 3 --frames-per-step 128 --distinct-images 64 --no-cpu-baseline --no-e2e > gpurun_out/r2_base_cfg3.json 2> gpurun_out/r2_base_cfg3.err || tail -c 600 gpurun_out/r2_base_cfg3.err
python bench.py --workload cfg1 --steps 4 --warmup 3 --distinct-images 64 --no-cpu-baseline --no-e2e > gpurun_out/r2_base_cfg1.json 2> gpurun_out/r2_base_cfg1.err || tail -c 600 gpurun_out/r2_base_cfg1.err
python - <<'PY'
import json
for f in ("cfg2","cfg3","cfg1"):
    try:
        d=json.loads(open(f"gpurun_out/r2_base_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "e2e", d["e2e"]["value"], {k:round(v["us_per_launch"],1) for k,v in d["roofline"]["kernels"].items()}, "upd/s %.3g"%d["voxel_updates_per_s"], "frac %.4f"%d["roofline"]["frac"], d["config"].get("chunk_retries"), d["config"].get("updates_per_frame"), d["config"].get("map_voxels_end"), d["config"].get("table_slots"))
    except Exception as e: print(f,"ERR",e)
PY

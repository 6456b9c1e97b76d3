#!/bin/bash
# Round-end style measurements on one B200: tests, smoke, bench (ours + reference arm), the other
# single-GPU workloads, ncu launch list and one full capture (warm caches: --cache-control none).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null
python bench.py --no-adaptive --no-cpu-baseline --no-e2e > gpurun_out/bench_noadapt.json 2>/dev/null
for wl in cfg1 cfg3; do
  fps=250; [ $wl = cfg3 ] && fps=64
  python bench.py --workload $wl --steps 4 --warmup 3 --frames-per-step $fps --distinct-images 64 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err || tail -3 gpurun_out/bench_$wl.err
done
CMD="python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --frames-per-step 64"
$CMD > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_l_final.log 2>&1
$CMD > gpurun_out/plain2_final.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_apply_chunk|k_expand" -s 12 -c 4 -o gpurun_out/prof_final -f $CMD > gpurun_out/ncu_f_final.log 2>&1
python - <<'PY'
import json
for f in ("bench_final","bench_ref_final","bench_noadapt","bench_cfg1","bench_cfg3"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), "e2e", round(d["e2e"]["value"]), "upd/s %.3g"%d["voxel_updates_per_s"], "frac", d.get("roofline",{}).get("frac"), d["config"].get("chunk_retries"), d.get("clocks"), d.get("gpu_launches"))
    except Exception as e: print(f,"ERR",e)
PY

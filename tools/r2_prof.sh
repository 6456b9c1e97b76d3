#!/bin/bash
# usage: tools/r2_prof.sh <tag> [workload] -- ncu launch list + one full capture of the two pipeline kernels
TAG=$1; WL=${2:-cfg2}
mkdir -p gpurun_out
FPS=64; [ $WL = cfg3 ] && FPS=32
CMD="python bench.py --workload $WL --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --frames-per-step $FPS --distinct-images 32"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_l_${TAG}.log 2>&1
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_apply_chunk|k_expand" -s 12 -c 4 -o gpurun_out/prof_${TAG} -f $CMD > gpurun_out/ncu_f_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_f_${TAG}.log
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r)>5 and r[0].isdigit()]
d=collections.defaultdict(list)
for r in rows:
    try: d[r[4].split("(")[0][:40]].append(float(r[-1]))
    except: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])): print(f"{k:42s} n={len(v):4d} mean={sum(v)/len(v)/1e3:9.1f} us total={sum(v)/1e6:8.2f} ms")
PY

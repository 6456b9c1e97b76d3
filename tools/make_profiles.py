#!/usr/bin/env python3
"""Turn the round-end captures in gpurun_out/ (tools/final_measure.sh) into the tracked summaries under profiles/.
usage: tools/make_profiles.py <round tag, e.g. r1>"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
CMD = "python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 --frames-per-step 64"

# 1. launch list (kept as captured) and per-kernel shares
rows = list(csv.reader(l for l in open(os.path.join(G, "launches_final.csv")) if not l.startswith("==")))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    mk = re.search(r"\b(k_\w+)", r[ki])
    name = mk.group(1) if mk else re.sub(r"^void ", "", r[ki]).split("<")[0].split("(")[0].strip()
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 200, same command as the bench (serialised, cold-cache replay: compare SHARES)\n# {CMD}   (cfg2, 16 frames per launch group)\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:44s} launches {n:4d}  total {us:9.1f} us  share {100 * us / tot:5.1f}%  avg {us / n:7.1f} us\n")
import shutil
shutil.copy(os.path.join(G, "launches_final.csv"), os.path.join(P, f"{tag}_launches_bench_cfg2.csv"))

# 2. full capture summary
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_report.py"), os.path.join(G, "prof_final.ncu-rep"),
                      "k_expand", "k_apply_chunk"], capture_output=True, text=True, env=dict(os.environ, NCU_TOP="12")).stdout
open(os.path.join(P, f"{tag}_ncu_full_summary.txt"), "w").write(out)

# 3. DRAM traffic per launch group vs algorithmic bytes
def metric(kern, name):
    m = re.search(r"## %s.*?\n((?:  .*\n)+)" % kern, out)
    for line in m.group(1).splitlines():
        if line.strip().startswith(name):
            return float(line.split()[1])
b = json.loads(open(os.path.join(G, "bench_final.json")).read().strip().splitlines()[-1])
H, W = 500, 512
alg = int(16 * (H * W + 24 * b["config"]["updates_per_frame"]))
tr = {"source": f"ncu --set full --clock-control none --cache-control none (warm L2, as in steady state), {CMD}, launches 12-15; profiles/{tag}_ncu_full_summary.txt",
      "frames_per_launch_group": 16}
total = 0
for k in ("k_expand", "k_apply_chunk"):
    r, w = metric(k, "dram__bytes_read.sum"), metric(k, "dram__bytes_write.sum")
    tr[k] = {"dram_bytes_read": int(r * 1e6), "dram_bytes_write": int(w * 1e6), "gpu_time_us": metric(k, "gpu__time_duration.sum")}
    total += int((r + w) * 1e6)
tr["dram_bytes_per_launch_group"] = total
tr["algorithmic_bytes_per_launch_group"] = alg
json.dump(tr, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_shares.txt")).read()); print(out[:3000]); print(json.dumps(tr, indent=1))

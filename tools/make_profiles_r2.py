#!/usr/bin/env python3
"""Turn the round-2 captures of tools/r2_profiles.sh (gpurun_out/) into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "sector": 1, "inst": 1}


def kname(s):
    return "k_apply_chunk" if "k_apply_chunk" in s else ("k_expand" if "k_expand" in s else s.split("(")[0][-40:])


def cold(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    d = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        d[kname(r[4])][r[-3]].append(float(r[-1].replace(",", "")) * UNIT.get(r[-2], 1))
    return {k: dict({m: sum(x) / len(x) for m, x in v.items()}, launches=len(next(iter(v.values())))) for k, v in d.items()}


def main():
    shutil.copyfile(os.path.join(G, "r2_launches_cfg2.csv"), os.path.join(P, "r2_launches_bench_cfg2.csv"))
    rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches_cfg2.csv"))) if len(r) > 5 and r[0].isdigit()]
    d = collections.defaultdict(list)
    for r in rows:
        d[kname(r[4])].append(float(r[-1].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(P, "r2_launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py cfg2, 64 frames per step (cold-cache, serialised:\n"
                "# compare shares, not absolutes).  The launch list contains the timed pass, the parity-free exclusive pass and set-up fills.\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:28s} launches={len(v):4d} mean={sum(v) / len(v) / 1e3:8.1f} us  share={100 * sum(v) / tot:5.1f} %\n")
    traffic = {}
    for cfg in ("cfg2", "cfg3"):
        c = cold(os.path.join(G, f"r2_cold_{cfg}.csv"))
        traffic[cfg] = {k: {"dram_bytes_per_launch": v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"],
                            "dram_read": v["dram__bytes_read.sum"], "dram_write": v["dram__bytes_write.sum"],
                            "us_per_launch_cold": v["gpu__time_duration.sum"] / 1e3, "l2_sectors": v["lts__t_sectors.sum"],
                            "warp_instructions": v["smsp__inst_executed.sum"], "launches_averaged": v["launches"]}
                        for k, v in c.items()}
    traffic["how"] = ("ncu --cache-control all (L2 flushed before every kernel replay) --clock-control none, "
                      "bench.py --workload cfgN --frames-per-step 64 (cfg2) / 32 (cfg3), launches 12.. of the run, 16 frames per launch; "
                      "k_expand's DRAM read therefore contains the 16 streamed frames")
    json.dump(traffic, open(os.path.join(P, "r2_traffic.json"), "w"), indent=1)
    rep = os.path.join(G, "r2_full_cfg2.ncu-rep")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_report.py"), rep, "k_expand", "k_apply_chunk"],
                         capture_output=True, text=True, env=dict(os.environ, NCU_TOP="14")).stdout
    with open(os.path.join(P, "r2_ncu_full_summary.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --cache-control none --import-source on, bench.py cfg2, 64 frames per step, 16 frames per launch\n" + out)
    shutil.copyfile(os.path.join(G, "r2_microbench_rmw.txt"), os.path.join(P, "r2_microbench_rmw.txt"))
    so = os.path.join(ROOT, "sonar_3d_reconstruction_b200", "csrc", "libsonar3d.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn, keep = None, []
    counts = collections.Counter()
    for ln in sass.splitlines():
        if "Function :" in ln:
            fn = ln.split("Function :")[1].strip()
        for op in ("UTMALDG", "SYNCS", "UBLKCP", "LDGSTS", "DADD", "DMUL", "FFMA", "ATOMS", "ATOMG", "RED."):
            if op in ln and fn and "k_expandIjLb0ELb0ELb0" in fn:
                counts[op] += 1
                if op in ("UTMALDG", "SYNCS") and len(keep) < 12:
                    keep.append(ln.strip())
    ptx = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                          "-shared", "-Xptxas", "-v", "-o", "/tmp/_s3d_ptxas.so", os.path.join(ROOT, "sonar_3d_reconstruction_b200", "csrc", "sonar3d.cu")],
                         capture_output=True, text=True).stderr
    want, grab = [], 0
    for ln in ptx.splitlines():
        if "Compiling entry function" in ln and ("k_expandIjLb0ELb0ELb0" in ln or "k_apply_chunkIjLb0" in ln):
            grab = 3
        if grab:
            want.append(ln.strip()); grab -= 1
    with open(os.path.join(P, "r2_sass_tma.txt"), "w") as f:
        f.write("# cuobjdump -sass csrc/libsonar3d.so, k_expand<u32, false, false, false>: opcode counts and the TMA / mbarrier instructions\n")
        f.write(json.dumps(dict(counts)) + "\n" + "\n".join(keep) + "\n\n# nvcc -Xptxas -v\n" + "\n".join(want) + "\n")
    print(open(os.path.join(P, "r2_launch_shares.txt")).read())
    print(json.dumps(traffic, indent=1)[:1500])
    print(open(os.path.join(P, "r2_sass_tma.txt")).read()[:1500])


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Summarise an .ncu-rep (read on the CPU box) into the text kept under profiles/.
usage: tools/ncu_report.py <report.ncu-rep> <kernel-regex> [<kernel-regex> ...]"""
import collections
import csv
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def source(rep, kern):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    lines, stalls, seen = {}, collections.Counter(), 0
    hdr = None
    for r in rows:
        if r and r[0] == "Function Name":
            seen += 1
            if seen > 1:
                break
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8:
            continue
        if r[0].isdigit():
            try:
                e = lines.setdefault(int(r[0]), [r[1], 0, 0])
                e[1] += int(r[6]); e[2] += int(r[7])
            except ValueError:
                pass
        elif r[2].startswith("0x"):
            for i, n in enumerate(hdr):
                if n.startswith("stall_") and "Not Issued" not in n and i < len(r) and r[i].isdigit():
                    stalls[n] += int(r[i])
    return lines, stalls


def main():
    rep, kerns = sys.argv[1], sys.argv[2:]
    hdr, units, rows = raw(rep)
    ki = hdr.index("Kernel Name")
    for kern in kerns:
        picked = [r for r in rows if re.search(kern, r[ki])]
        if not picked:
            continue
        print(f"## {kern}  ({len(picked)} captured launches; first shown)")
        r = picked[0]
        for mname in METRICS:
            if mname in hdr:
                print(f"  {mname:72s} {r[hdr.index(mname)]:>16s} {units[hdr.index(mname)]}")
        lines, stalls = source(rep, kern)
        tot = sum(v[2] for v in lines.values()) or 1
        samp = sum(stalls.values()) or 1
        print("  warp stall samples: " + ", ".join(f"{k[6:]} {100 * v / samp:.0f}%" for k, v in stalls.most_common(6)))
        print("  hottest source lines (share of executed warp instructions / stall samples):")
        for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][2])[:int(__import__("os").environ.get("NCU_TOP", "10"))]:
            print(f"    {ln:5d} {100 * v[2] / tot:5.1f}% {v[1]:6d}  {v[0][:96]}")
        print()


if __name__ == "__main__":
    main()

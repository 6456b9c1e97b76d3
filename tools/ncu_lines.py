#!/usr/bin/env python3
"""Per-source-line executed warp instructions and stall samples of one kernel in an .ncu-rep, all lines, in file order.
usage: tools/ncu_lines.py <report> <kernel-regex> [min_pct]"""
import sys
sys.path.insert(0, __import__("os").path.dirname(__file__))
from ncu_report import source
rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
lines, stalls = source(rep, kern)
tot = sum(v[2] for v in lines.values()) or 1
st = sum(v[1] for v in lines.values()) or 1
for ln in sorted(lines):
    v = lines[ln]
    if 100 * v[2] / tot >= minp or 100 * v[1] / st >= minp:
        print(f"{ln:5d} {100 * v[2] / tot:5.1f}% inst {100 * v[1] / st:5.1f}% stall  {v[0][:110]}")
print("total warp inst", tot, "stall samples", st)

#!/usr/bin/env python3
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export by CUDA source line:
share of stall samples, warp-instructions and average active threads per line. Usage: ncu_lines.py export.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ei, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try:
            s, e, t = int(r[si] or 0), int(r[ei] or 0), int(r[ti] or 0)
        except ValueError:
            continue
        if s or e:
            k = (cur_file, int(r[0]))
            a = agg.get(k, (0, 0, 0, r[1][:100]))
            agg[k] = (a[0] + s, a[1] + e, a[2] + t, a[3])
tot = sum(v[0] for v in agg.values())
tote = sum(v[1] for v in agg.values())
print("total samples", tot, "warp-instructions", tote)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-18s %4d  smp %5.1f%%  inst %5.1f%%  thr/inst %5.1f  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / max(1, tote), v[2] / max(1, v[1]), v[3].strip()))

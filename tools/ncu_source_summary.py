#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export: share of samples/instructions in the hot loop vs the rest,
stall mix of each, and the hottest instructions outside the loop. Usage: ncu_source_summary.py src.csv [kernel-index]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
# the export holds one table per profiled launch: "Kernel Name" row, header row, data rows
tables, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        tables.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
t = tables[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
hdr, data = t["hdr"], t["data"]
print(t["name"][:100], "instructions:", len(data))
ai, si, ni, ei, ti = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
I = lambda r, i: int(r[i] or 0)
tot, tote = sum(I(r, ni) for r in data), sum(I(r, ei) for r in data)
mx = max(I(r, ei) for r in data)
hot = [r for r in data if I(r, ei) > 0.9 * mx]
hotset = set(id(r) for r in hot)
rest = [r for r in data if id(r) not in hotset]
hs, he = sum(I(r, ni) for r in hot), sum(I(r, ei) for r in hot)
print("hot loop: %d instrs, executed %d each; samples %.1f%%, warp-instructions %.1f%%" % (len(hot), mx, 100 * hs / tot, 100 * he / tote))
for k in stalls:
    i = hdr.index(k)
    a, b = sum(I(r, i) for r in hot), sum(I(r, i) for r in rest)
    if a + b > 0.005 * tot:
        print("  %-26s hot %5.1f%%   rest %5.1f%%" % (k, 100 * a / max(1, hs), 100 * b / max(1, tot - hs)))
ex = [r for r in rest if I(r, ei) > 0]
print("rest: avg active threads per warp-instruction %.2f; fp64-ish share n/a" % (sum(I(r, ti) for r in ex) / max(1, sum(I(r, ei) for r in ex))))
ex.sort(key=lambda r: -I(r, ni))
for r in ex[:30]:
    print("  %s %-58s samples %6d  exec %10d  thr/inst %.1f" % (r[ai][-5:], r[si][:58], I(r, ni), I(r, ei), I(r, ti) / max(1, I(r, ei))))

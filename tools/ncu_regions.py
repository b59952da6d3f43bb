#!/usr/bin/env python3
"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export by enclosing C++ function:
stall samples, warp-instructions, active threads per instruction, distinct SASS instructions (total / executed >= 1M times,
i.e. the instruction-cache footprint of the code that runs once per ray segment), plus the kernel-wide stall mix.

    python tools/ncu_regions.py gpurun_out/src_lines.csv [gpurun_out/src.csv]
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tray_b200", "csrc")
FUNC = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__forceinline__|__noinline__|\s)+[\w:<>,\s\*&]*?\b(\w+)\s*\(")


def function_starts(path):
    """[(line, name)] of the functions defined in a source file (good enough for these headers)."""
    out = []
    try:
        lines = open(path).read().split("\n")
    except OSError:
        return out
    for n, l in enumerate(lines, 1):
        if "__device__" in l or "__global__" in l:
            m = re.search(r"\b(\w+)\s*\((?!.*\)\s*;)", l.split("//")[0])
            names = re.findall(r"\b([A-Za-z_]\w*)\s*\(", l.split("//")[0])
            names = [x for x in names if x not in ("__launch_bounds__", "__align__", "sizeof", "if", "for", "while", "return")]
            if names:
                out.append((n, names[0]))
        elif re.match(r"^\s*auto (\w+) = \[", l):
            out.append((n, "  lambda " + re.match(r"^\s*auto (\w+) = \[", l).group(1)))
    return out


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    starts = {}
    cur_file, line, agg = None, None, collections.OrderedDict()
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            if cur_file not in starts:
                starts[cur_file] = function_starts(os.path.join(CSRC, cur_file))
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0].isdigit():
            line = int(r[0])
            continue
        if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
            try:
                e, s = int(r[7]), int(r[6])
            except ValueError:
                continue
            fn = cur_file
            for ln, name in starts.get(cur_file, []):
                if ln <= line:
                    fn = name
            a = agg.setdefault((cur_file, fn), [0, 0, 0, 0])
            a[0] += 1
            a[1] += 1 if e >= 1000000 else 0
            a[2] += s
            a[3] += e
    tots = sum(a[2] for a in agg.values()) or 1
    tote = sum(a[3] for a in agg.values()) or 1
    print("%-22s %-30s %7s %7s %7s %7s" % ("file", "function", "instrs", "hot", "smp%", "exec%"))
    for (f, fn), a in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        if a[2] > 0.002 * tots or a[1] > 10:
            print("%-22s %-30s %7d %7d %6.1f%% %6.1f%%" % (f, fn[:30], a[0], a[1], 100 * a[2] / tots, 100 * a[3] / tote))
    print("total distinct SASS instructions %d, executed >= 1M times: %d (%.1f KB)" % (
        sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values()), sum(a[1] for a in agg.values()) * 16 / 1024))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
        hdr = rows[hi[0] + 1]
        data = [r for r in rows[hi[0] + 2:] if len(r) == len(hdr)]
        ei, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        I = lambda r, i: int(r[i] or 0)
        tot = sum(I(r, si) for r in data) or 1
        print("kernel:", rows[hi[0]][1][:90])
        print("warp-instructions %d, active threads per instruction %.2f" % (sum(I(r, ei) for r in data), sum(I(r, ti) for r in data) / max(1, sum(I(r, ei) for r in data))))
        for k in [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]:
            v = sum(I(r, hdr.index(k)) for r in data)
            if v > 0.01 * tot:
                print("  %-26s %5.1f%%" % (k, 100 * v / tot))


if __name__ == "__main__":
    main()

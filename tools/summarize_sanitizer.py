#!/usr/bin/env python3
"""gpurun_out/sanitizer_<tool>.log (tools/gpu_sanitize.sh) -> profiles/<tag>_sanitizer.txt: the summary lines of every
compute-sanitizer tool plus the first reports, if any. Usage: python tools/summarize_sanitizer.py r02"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = ["compute-sanitizer (CUDA 12.9) over tools/sanitize_run.py on one B200: every trace-kernel variant (plain / regroup / wavefront layouts x",
       "linear / BVH / cluster structures x four arithmetic modes), cluster tables of every size remainder (n = 0..20, 63..65, 127, 129, 511, 513),",
       "device LBVH build, progressive sums, present, PNG encoder, conformance kernel, MaxDepth 0, generator variants.", ""]
for tool in ("memcheck", "racecheck", "initcheck", "synccheck"):
    p = os.path.join(ROOT, "gpurun_out", "sanitizer_%s.log" % tool)
    if not os.path.exists(p):
        out.append("%s: not run" % tool)
        continue
    txt = open(p, errors="replace").read()
    lines = txt.splitlines()
    summ = [l for l in lines if re.search(r"ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok|Internal Sanitizer Error|Target application", l)]
    out.append("== %s ==" % tool)
    out += ["  " + l.strip() for l in summ[-4:]] or ["  (no summary line)"]
    reports = [l for l in lines if l.startswith("=========") and re.search(r"Invalid|Race|Uninitialized|hazard|Barrier error|Error:", l)]
    for l in reports[:12]:
        out.append("  " + l.strip())
    out.append("")
open(os.path.join(ROOT, "profiles", tag + "_sanitizer.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))

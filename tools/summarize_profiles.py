#!/usr/bin/env python3
"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (named per round).
Usage: python tools/summarize_profiles.py r01"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list (gpu__time_duration.sum per launch) -> per-kernel totals and shares
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
with open(os.path.join(P, tag + "_launches.csv"), "w") as f:
    f.write("launch,kernel,duration_ms\n")
    for n, r in enumerate(rows[1:]):
        v = float(r[vi].replace(",", ""))
        ms = v / 1e6 if r[ui].startswith("n") else (v / 1e3 if r[ui].startswith("u") else v)
        name = r[ki].split("(")[0][:90]
        f.write("%d,%s,%.4f\n" % (n, name.replace(",", ";"), ms))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, tag + "_launch_summary.txt"), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none: `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt`\n")
    f.write("(cold-cache, serialised: compare shares, not absolutes)\n\n%-70s %6s %12s %7s\n" % ("kernel", "n", "total ms", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-70s %6d %12.3f %6.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))

# 2. full capture of the trace kernel -> selected raw metrics
which = sys.argv[2] if len(sys.argv) > 2 else "prof_trace"
suffix = "" if which == "prof_trace" else "_" + which.replace("prof_trace_", "")
rep = os.path.join(G, which + ".ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum")
with open(os.path.join(P, tag + "_trace_kernel%s_metrics.csv" % suffix), "w") as f:
    f.write("launch_id,kernel,metric,value,unit\n")
    for r in rr[2:]:
        for i, name in enumerate(h):
            if name in keep or (name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio")):
                f.write("%s,%s,%s,%s,%s\n" % (r[h.index("ID")], r[h.index("Kernel Name")].split("(")[0].replace(",", ";"), name, r[i].replace(",", ""), u[i]))

# 3. source page -> hot loop vs rest
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tmp = os.path.join(G, "src.csv")
open(tmp, "w").write(src)
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_summary.py"), tmp], capture_output=True, text=True).stdout
src2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp2 = os.path.join(G, "src_lines.csv")
open(tmp2, "w").write(src2)
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp2, "45"], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_trace_kernel%s_source_summary.txt" % suffix), "w").write(
    "ncu --set full --import-source on, source page of tray::trace_kernel (first profiled launch)\n\n" + out +
    "\nby CUDA source line (stall samples, warp-instructions, active threads per instruction):\n" + lines)
print(open(os.path.join(P, tag + "_launch_summary.txt")).read())
print(out[:1500])

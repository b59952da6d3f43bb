#!/usr/bin/env python3
"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (named per round).
Usage: python tools/summarize_profiles.py r01"""
import collections, csv, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list (gpu__time_duration.sum per launch) -> per-kernel totals and shares
try:
    rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10]
except OSError:
    rows = []
hdr = rows[0] if rows else ["Kernel Name", "Metric Value", "Metric Unit"]
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
        f.write("%-70s %6d %12.3f %6.1f%%\n" % (k, a[0], a[1], 100 * a[1] / max(tot, 1e-9)))

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

# 3. source page -> per-function table, per-line table, and the JSON summary bench.py reads (roofline.fp64_pipe_busy, ...)
import json
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
tmp = os.path.join(G, "src.csv")
open(tmp, "w").write(src)
src2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp2 = os.path.join(G, "src_lines.csv")
open(tmp2, "w").write(src2)
regions = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), tmp2, tmp], capture_output=True, text=True).stdout
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), tmp2, "45"], capture_output=True, text=True).stdout
open(os.path.join(P, tag + "_trace_kernel%s_source_summary.txt" % suffix), "w").write(
    "ncu --set full --import-source on, source page of tray::trace_kernel (first profiled launch)\n\n"
    "by enclosing function (distinct SASS instructions, of which executed >= 1M times, stall samples, warp-instructions):\n" + regions +
    "\nby CUDA source line (stall samples, warp-instructions, active threads per instruction):\n" + lines)
# time shares by what the code is for (stall samples of the function table)
GROUPS = {
    "scan_boxes_and_prefilter": ("cluster_scan", "lambda boxes", "lambda pair", "lambda may_hit8", "lambda lds4", "lambda rcp", "sm_100_rt.hpp",
                                 "sm_80_rt.hpp", "sm_32_intrinsics.hpp", "filter_scan", "rcp_f32", "lambda merge4", "lambda ldg4", "cluster_scan_big"),
    "exact_resolve": ("resolve_candidates_lex", "resolve_candidates", "sphere_terms", "flush_candidates_lex", "miss_bits", "certainly_missed", "push_candidates"),
    "ieee_div_sqrt_fp64": ("div3_f64", "sqrt_f64", "div_f64", "tsqrt", "tdiv", "div3_body", "div3_hot", "rcp_refined", "div_tail", "div_tail_ok",
                           "div_by_rcp", "tsqrt_hot", "trcp", "tdiv_r", "unit"),
    "generators": ("pcg_step", "pcg_u64", "pcg_f64", "pcg_norm", "pcg_unit_vector", "go_log", "go_exp", "pcg_step_body", "pcg_u64_inline", "pcg_f64_pair"),
    "regenerate_camera_rays": ("camera_sample", "get_ray", "get_ray_drawn", "pcg_in_disc"),
}
share = {k: 0.0 for k in GROUPS}
share["shade_and_loop_body"] = 0.0
for l in regions.splitlines():
    m = l.split()
    if len(m) >= 6 and m[-1].endswith("%") and m[-2].endswith("%"):
        fn = " ".join(m[1:-4])
        try:
            smp = float(m[-2].rstrip("%")) / 100
        except ValueError:
            continue
        for g, names in GROUPS.items():
            if fn in names or m[0] in names:
                share[g] += smp
                break
        else:
            share["shade_and_loop_body"] += smp
def metric(name):
    for r in rr[2:3]:
        if name in h:
            try:
                return float(r[h.index(name)].replace(",", ""))
            except ValueError:
                return None
    return None
stalls = {}
for name in h:
    if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio"):
        v = metric(name)
        if v and v > 0.05:
            stalls[name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = v
summary = {
    "kernel": rr[2][h.index("Kernel Name")] if len(rr) > 2 else None,
    "duration_ms_under_ncu": (metric("gpu__time_duration.sum") or 0) / (1e6 if u[h.index("gpu__time_duration.sum")].startswith("n") else 1.0),
    "registers": metric("launch__registers_per_thread"),
    "fp64_pipe_busy": (metric("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or 0) / 100,
    "fp32_pipe_slot_frac": (metric("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or 0) / 100,
    "alu_pipe_frac": (metric("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") or 0) / 100,
    "issue_active": (metric("smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100,
    "threads_per_instruction": metric("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "warps_active_frac": (metric("sm__warps_active.avg.pct_of_peak_sustained_active") or 0) / 100,
    "warp_instructions": metric("smsp__inst_executed.sum"),
    "dram_bytes": (metric("dram__bytes_read.sum") or 0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[h.index("dram__bytes_read.sum")], 1) +
                  (metric("dram__bytes_write.sum") or 0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u[h.index("dram__bytes_write.sum")], 1),
    "stalled_warps_per_issue": stalls,
    "time_share": share,
    "how": "ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-alt",
}
if which == "prof_trace":
    json.dump(summary, open(os.path.join(P, tag + "_trace_kernel_summary.json"), "w"), indent=1)
try:
    print(open(os.path.join(P, tag + "_launch_summary.txt")).read())
except OSError:
    pass
print(regions[:3000])
print(json.dumps(summary, indent=1))

#!/usr/bin/env python3
"""bench.py -- throughput of the tray path-tracing hot path on B200 (Mpaths/s), driver contract.

    python bench.py --gpus 1 --steps K --warmup W                      # own arm, 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # N GPUs, tile sharding
    python bench.py --impl reference ...                               # reference arm: CPU port on host cores

A "step" is one full render of the workload (one pass of the hot path over every pixel x sample).
N=1 workload: BASELINE.json configs[1] (RichScene seed 2, 1920x1080, 64 rays/pixel, depth 50, fp64).
N>1 workload: configs[2] (3840x2160, 256 rays/pixel, depth 50), interleaved rows across ranks,
no data-path collective (each rank writes its own rows of the shared host image).

Timed regions:
  value  -- device-resident: scene already in HBM, image left in HBM; CUDA events recorded by the library
            on its launch stream around each step's kernels, summed over K steps, MAX over ranks.
  e2e    -- the reference-facing call sequence with HOST buffers every step: tray_scene_upload (H2D of the
            flattened scene) + tray_render into a host RGBA buffer (D2H inside the call); wall clock
            bracketed by synchronize + barrier, MAX over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (width, height, spp, depth, grid half-width, description)
    "config1": (400, 225, 10, 50, 11, "benchmark scene seed 2, 400x225, 10 rays/pixel, depth 50"),
    "config2": (1920, 1080, 64, 50, 11, "benchmark scene seed 2, 1920x1080, 64 rays/pixel, depth 50"),
    "config3": (3840, 2160, 256, 50, 11, "benchmark scene seed 2, 3840x2160, 256 rays/pixel, depth 50"),
    "config4": (1920, 1080, 64, 12, 50, "synthetic dense scene (~10k spheres), 1920x1080, 64 rays/pixel, depth 12"),
    "config5": (640, 360, 64, 12, 11, "interactive tray path: 160x45 terminal, -s 4 -> 640x360, 64 rays/pixel, depth 12"),
}
SEED = 2
FLOPS_PER_TEST = 18.0     # miss path of Sphere.Hit, `a` hoisted
FLOPS_PER_SEGMENT = 155.0  # a = D.D (5) + closest-hit finish, scatter / sky, RNG fp (150)


def algorithmic_flops(segments, n_spheres, sphere_tests=None):
    """18 flops per Sphere.Hit evaluated + 155 per segment; N_tested = N for the linear scan, kernel-counted for the BVH."""
    tests = float(segments) * n_spheres if sphere_tests is None else float(sphere_tests)
    return FLOPS_PER_TEST * tests + FLOPS_PER_SEGMENT * float(segments)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one trace_kernel launch from the committed `ncu --set full`
    capture (profiles/, same command line); None when no capture is committed."""
    import csv, glob
    files = sorted(glob.glob(os.path.join(_ROOT, "profiles", "r*_trace_kernel_metrics.csv")))
    if not files:
        return None, None
    tot, first = 0.0, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for r in csv.DictReader(open(files[-1])):
        first = first if first is not None else r["launch_id"]
        if r["launch_id"] == first and r["metric"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r["value"]) * scale.get(r["unit"], 1.0)
    return (tot or None), os.path.basename(files[-1])


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_port_run(workload, threads, target_seconds, fma_mode=0):
    """Times the oracle (a C port of the reference's CPU loop) on a bounded sample of the workload:
    rows y == 0 (mod step) of the image, per-sample streams, all host threads."""
    from oracle import oracle as O
    w, h, spp, depth, half, _ = WORKLOADS[workload]
    scene = O.rich_scene(SEED, half)
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    p = O.make_params(w, h, spp=spp, max_depth=depth, seed=SEED, num_workers=threads, stream_mode=1, fma_mode=fma_mode)
    # calibrate on a few rows spread over the image, then size the sample for ~target_seconds
    cal_step = max(1, h // 4)
    t0 = time.perf_counter()
    _, st = O.render_sampled_rows(scene, cam, p, cal_step, cal_step // 2, threads)
    dt = time.perf_counter() - t0
    rate = st["paths"] / max(dt, 1e-6)
    rows = int(max(threads, min(h, target_seconds * rate / (w * spp))))
    step = max(1, h // rows)
    return dict(scene=scene, cam=cam, params=p, step=step, threads=threads, rows=len(range(0, h, step)), O=O, w=w, h=h, spp=spp)


def cpu_port_step(cp):
    O = cp["O"]
    t0 = time.perf_counter()
    img, st = O.render_sampled_rows(cp["scene"], cp["cam"], cp["params"], cp["step"], 0, cp["threads"])
    dt = time.perf_counter() - t0
    return img, st, dt


def oracle_rows_check(workload, host_img, n_rows, threads):
    """Parity spot check: `n_rows` rows spread over the image (y == off mod step) of `workload` rendered by the CPU oracle
    (strict fp64, per-sample streams, one row per host thread at a time) and compared with the same rows of the GPU image.
    Returns (parity dict, cpu stats)."""
    from oracle import oracle as O
    w, h, spp, depth, half, _ = WORKLOADS[workload]
    scene = O.rich_scene(SEED, half)
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    p = O.make_params(w, h, spp=spp, max_depth=depth, seed=SEED, num_workers=threads, stream_mode=1, fma_mode=0)
    step = max(1, h // max(1, min(h, n_rows)))
    off = step // 2
    t0 = time.perf_counter()
    img, st = O.render_sampled_rows(scene, cam, p, step, off, threads)
    dt = time.perf_counter() - t0
    ys = list(range(off, h, step))
    d = np.abs(img[ys, :, :3].astype(np.int16) - host_img[ys, :, :3].astype(np.int16)).max(axis=2)
    parity = {"rows_checked": len(ys), "pixels_identical_frac": float((d == 0).mean()), "pixels_within_1lsb_frac": float((d <= 1).mean()),
              "mode": "GPU fp64 vs strict CPU port (oracle), rows y%%%d==%d" % (step, off)}
    return parity, {"paths": st["paths"], "seconds": dt, "threads": threads}


def ncu_summary():
    """Pipe utilisation and where the time goes, from the committed `ncu --set full` capture of the default kernel
    (profiles/r*_trace_kernel_summary.json, written by tools/summarize_profiles.py); None when there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(_ROOT, "profiles", "r*_trace_kernel_summary.json")))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
        d["source"] = os.path.basename(files[-1])
        return d
    except Exception:
        return None


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; Go cannot be built here) on all host
    threads, same metric/config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    workload = args.workload or ("config2" if args.gpus == 1 else "config3")
    threads = host_threads()
    total_budget = args.ref_budget
    per_step = max(0.5, min(25.0, total_budget / max(1, args.steps + args.warmup)))
    cp = cpu_port_run(workload, threads, per_step)
    for _ in range(args.warmup):
        cpu_port_step(cp)
    paths, segs, secs = 0, 0, 0.0
    for _ in range(args.steps):
        _, st, dt = cpu_port_step(cp)
        paths += st["paths"]; segs += st["segments"]; secs += dt
    w, h, spp, depth, half, desc = WORKLOADS[workload]
    val = paths / secs / 1e6
    sample = "rows y%%%d==0 of %s (%d rows, %.2f Mpaths/step), per-sample streams, strict fp64 (Go/amd64 semantics)" % (
        cp["step"], workload, cp["rows"], paths / args.steps / 1e6)
    line = {"impl": "reference", "metric": "Mpaths/s", "value": val, "unit": "Mpaths/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "mrays_per_s": segs / secs / 1e6,
            "config": {"workload": workload, "desc": desc, "width": w, "height": h, "rays_per_pixel": spp, "max_depth": depth,
                       "seed": SEED, "spheres": cp["scene"].n},
            "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "Go toolchain absent: the reference arm is the C oracle port of ray/tracer.go's loop on all host threads"}
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tray_b200 import ray, rand, _lib, multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run --nproc-per-node %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tray_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side barriers: ranks that wait must not keep a kernel spinning on their GPU

    workload = args.workload or ("config2" if args.gpus == 1 else "config3")
    w, h, spp, depth, half, desc = WORKLOADS[workload]
    PREC = {"fp64": ray.FP64_STRICT, "fp64-brute": ray.FP64_STRICT_BRUTE, "fp64-fma": ray.FP64_FMA, "fp32": ray.FP32}
    precision = PREC[args.precision]
    LAYOUTS = {"auto": ray.LAYOUT_AUTO, "plain": ray.LAYOUT_PLAIN, "regroup": ray.LAYOUT_REGROUP, "wavefront": ray.LAYOUT_WAVEFRONT}

    ctx = ray.Context([local_rank])

    def job(name, shard=(0, 0)):
        """Scene, camera and params of a workload on this rank's context (Tracer.Render's defaulting: ray/tracer.go:49-83)."""
        jw, jh, jspp, jdepth, jhalf, _ = WORKLOADS[name]
        sc = ray.RichScene(rand.New(SEED), jhalf)
        t = ray.New(jw, jh)
        t.Camera = ray.RichSceneCamera()
        t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision = jdepth, jspp, SEED, precision
        t.Layout = LAYOUTS[args.layout]
        t.ShardIndex, t.ShardCount = shard
        t.Context = ctx
        t._prepare(sc)
        return t, sc.flatten()

    split_samples = args.split == "samples" and world > 1
    tr, flat = job(workload, (rank, world) if (world > 1 and not split_samples) else (0, 0))
    n_spheres = len(flat["cx"])
    ctx.upload(flat)
    cam_c = tr.to_c()
    params = tr._params(0, h)

    def sample_params():
        q = tr._params(0, h)
        q.shard_index, q.shard_count = 0, 0
        q.sample_offset, q.sample_stride, q.sample_count = multi.sample_subset(spp, rank, world)
        q.sums_mode = ray.SUMS_OVERWRITE
        return q

    if split_samples:  # rank r traces samples s == r (mod world) of every pixel; sums reduced to rank 0 with NCCL
        params = sample_params()

    # shared host image for the e2e leg (each rank writes its own row bands; no collective)
    if world > 1:
        shm = "/dev/shm/tray_bench_%s.rgba" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            np.lib.format.open_memmap(shm, mode="w+", dtype=np.uint8, shape=(h, w, 4)).flush()
        dist.barrier()
        host_img = np.load(shm, mmap_mode="r+")
    else:
        host_img = np.zeros((h, w, 4), dtype=np.uint8)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def host_barrier():
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=cpu_group)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    sums_t = [None]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def exchange(out_img=None):
        """Sample split: the one exchange step -- sum-reduce of the fp64 colour sums to rank 0 (NCCL over NVLink), then
        1/N + sRGB on rank 0. Returns the device time of the reduce on torch's stream (CUDA events)."""
        ptr, n = ctx.device_sums()
        if sums_t[0] is None or sums_t[0][0] != (ptr, n):
            sums_t[0] = ((ptr, n), torch.as_tensor(multi._DeviceBuffer(ptr, n), device=torch.device("cuda", local_rank)))
        ev[0].record()
        dist.reduce(sums_t[0][1], dst=0, op=dist.ReduceOp.SUM)
        ev[1].record()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if rank == 0:
            ctx.resolve_sums(spp, out_img)
        return ev[0].elapsed_time(ev[1]) + (time.perf_counter() - t0) * 1e3

    peak_tf, _ = ctx.measure_peak(0)  # DFMA issue-bound peak of this GPU, measured live
    peak_strict_tf, _ = ctx.measure_peak(1)
    peak_f32_tf, _ = ctx.measure_peak(2)

    def timed_steps(cam, prm, steps, with_exchange=False):
        """`steps` device-resident renders (scene in HBM, image left in HBM), L2 flushed in between; library CUDA events."""
        acc = dict(ms=0.0, trace_ms=0.0, launches=0, segments=0, paths=0, trace_launches=0, sphere_tests=0, box_tests=0.0, exchange_ms=0.0)
        for _ in range(steps):
            flush_buf.zero_()  # L2 flush between timed iterations (outside the event-timed kernels)
            torch.cuda.synchronize()
            st = ctx.render(cam, prm, None)
            acc["ms"] += st["kernel_ms"]; acc["trace_ms"] += st["trace_kernel_ms"]; acc["launches"] += st["launches"]
            if with_exchange:
                xms = exchange()
                acc["ms"] += xms; acc["exchange_ms"] += xms; acc["launches"] += 1
            acc["segments"] += st["segments"]; acc["paths"] += st["paths"]; acc["trace_launches"] += int(st["trace_launches"])
            acc["sphere_tests"] += st["sphere_tests"]; acc["box_tests"] += st["box_tests"]
        return acc

    # ---- device-resident leg ----
    for _ in range(max(args.warmup, 3)):  # timing rules: at least 3 warm-up steps
        ctx.render(cam_c, params, None)
        if split_samples:
            exchange()
    sampler = ClockSampler(local_rank)
    sync_all()
    sampler.start()
    wall0 = time.perf_counter()
    m = timed_steps(cam_c, params, args.steps, split_samples)
    sync_all()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    dev_ms, trace_ms, launches, segments, paths = m["ms"], m["trace_ms"], m["launches"], m["segments"], m["paths"]
    trace_launches, sphere_tests, box_tests, exchange_ms = m["trace_launches"], m["sphere_tests"], m["box_tests"], m["exchange_ms"]
    per_rank = None
    if world > 1:  # every rank's own device time (ms per step: all kernels | trace kernel): where the max over ranks comes from
        t = torch.tensor([dev_ms / args.steps, trace_ms / args.steps], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        per_rank = {"ms_per_step": [round(float(x[0]), 3) for x in g], "trace_kernel_ms": [round(float(x[1]), 3) for x in g]}
    dev_ms = max_over_ranks(dev_ms)
    all_paths = sum_over_ranks(paths)
    all_segments = sum_over_ranks(segments)
    value = all_paths / (dev_ms * 1e-3) / 1e6
    # roofline of the dominant kernel (trace_kernel) on this rank, SURVEY 8(d): sum over segments of (18 * N_tested + 155) flops, with
    # "N_tested = N for brute force and = kernel-counted tests for BVH runs". The default closest hit is a hierarchy (boxes over
    # chunks of spheres), i.e. a BVH run in that sense: `achieved` / `frac` count the sphere tests the kernel made (a few per cent of
    # N). The figure with the reference algorithm's N tests per Scene.Hit -- reference-equivalent work per second, which exceeds
    # the DFMA peak -- is reported beside it as frac_reference_equivalent_work.
    structure = "linear scan" if sphere_tests >= segments * n_spheres else ("two-level clusters" if box_tests > 0 else "bvh")
    flops_reference = algorithmic_flops(segments, n_spheres)
    flops = flops_reference if structure == "linear scan" else algorithmic_flops(segments, n_spheres, sphere_tests)
    achieved_tf = flops / (trace_ms * 1e-3) / 1e12 if trace_ms > 0 else 0.0
    peak_used = {"fp64": peak_tf, "fp64-brute": peak_tf, "fp64-fma": peak_tf, "fp32": peak_f32_tf}[args.precision]
    peak_ffma2_tf = ctx.measure_peak(5)[0]
    try:
        loop_probe_tf = ctx.measure_peak({"fp64": 4, "fp64-brute": 4, "fp64-fma": 3, "fp32": 3}[args.precision])[0]
    except Exception:
        loop_probe_tf = None  # table larger than shared memory (config 4)

    # ---- the other fp64 kernels, reported beside the default (same steps, device-resident) ----
    alts = []
    frac_fp64_brute = None
    if args.precision == "fp64" and not args.no_alt and not split_samples:
        notes = {"fp64-brute": "same strict arithmetic, every test on the FP64 pipe (no pre-filter, linear scan): 17 FP64 instructions per 18-flop test, "
                               "structural ceiling 0.529 of the DFMA peak -- the kernel that actually sits on the FP64 roofline",
                 "fp64-fma": "Sphere.Hit discriminant with fused multiply-add (11 FP64 instructions/test, ceiling 0.818), no pre-filter; image "
                             "identical at 8 bits, first-hit ids exact, t within 3e-12, normals within 4e-11 of the strict result"}
        for name in ("fp64-brute", "fp64-fma"):
            p2 = tr._params(0, h)
            p2.precision = PREC[name]
            for _ in range(2):
                ctx.render(cam_c, p2, None)
            a = timed_steps(cam_c, p2, args.steps)
            a_ms = max_over_ranks(a["ms"])
            a_tf = algorithmic_flops(a["segments"], n_spheres) / (a["trace_ms"] * 1e-3) / 1e12
            if name == "fp64-brute":
                frac_fp64_brute = a_tf / peak_tf
            alts.append({"precision": name, "value": sum_over_ranks(a["paths"]) / (a_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
                         "roofline": {"bound": "fp64", "achieved": a_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": a_tf / peak_tf},
                         "note": notes[name]})
        # same default arithmetic, other divergence layout / closest-hit structure (results bit-identical)
        for name, setp, note in (
                ("fp64, regroup layout", lambda q: setattr(q, "layout", ray.LAYOUT_REGROUP),
                 "default kernel plus per-material regrouping of the CTA's paths through shared memory after every Scene.Hit (TRAY_LAYOUT_REGROUP)"),
                ("fp64, wavefront layout", lambda q: setattr(q, "layout", ray.LAYOUT_WAVEFRONT),
                 "path state in HBM, one bounce = intersect | shade over per-material queues | regenerate kernels (TRAY_LAYOUT_WAVEFRONT; linear pre-filter scan)"),
                ("fp64, linear scan", lambda q: setattr(q, "accel", ray.ACCEL_BRUTE),
                 "round-1 default: every sphere of the table through the exact fp32 pair pre-filter, reference order (TRAY_ACCEL_BRUTE)"),
                ("fp64, bvh", lambda q: setattr(q, "accel", ray.ACCEL_BVH),
                 "per-lane BVH traversal instead of the warp-wide cluster boxes (TRAY_ACCEL_BVH); image bit-identical")):
            p2 = tr._params(0, h)
            setp(p2)
            for _ in range(2):
                ctx.render(cam_c, p2, None)
            a = timed_steps(cam_c, p2, args.steps)
            a_ms = max_over_ranks(a["ms"])
            alts.append({"precision": name, "value": sum_over_ranks(a["paths"]) / (a_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
                         "sphere_tests_per_segment": a["sphere_tests"] / max(1, a["segments"]), "note": note})
        # fp32 fast path: throughput + PSNR of its 8-bit image against the fp64 image (same streams)
        if world == 1:
            p3 = tr._params(0, h)
            p3.precision = ray.FP32
            img64 = np.zeros((h, w, 4), dtype=np.uint8)
            img32 = np.zeros((h, w, 4), dtype=np.uint8)
            ctx.render(cam_c, params, img64)
            ctx.render(cam_c, p3, img32)
            f = timed_steps(cam_c, p3, args.steps)
            mse = float(((img64[:, :, :3].astype(np.float64) - img32[:, :, :3].astype(np.float64)) ** 2).mean())
            alts.append({"precision": "fp32", "value": f["paths"] / (f["ms"] * 1e-3) / 1e6, "unit": "Mpaths/s",
                         "psnr_db_vs_fp64": (10 * np.log10(255.0 ** 2 / mse)) if mse > 0 else None,
                         "pixels_within_1lsb_frac": float((np.abs(img64.astype(np.int16) - img32.astype(np.int16)).max(axis=2) <= 1).mean()),
                         "segments_per_path": f["segments"] / max(1, f["paths"]),
                         "note": "fast path: float32 arithmetic end to end on the same RNG streams (FrontEpsilon 1e-3; the sphere a ray leaves outwards is not "
                                 "re-tested; attenuation carried forward; 80 registers x 24 warps/SM); no parity claim, PSNR against the fp64 image"})

    # ---- end-to-end leg: host buffers, scene H2D + image D2H inside the timed region ----
    term_cols, term_rows = 160, 45

    def e2e_run(cam, prm, fl, img, steps, warm, interactive=False, with_exchange=False):
        """The reference-facing call sequence with HOST buffers every step. interactive: tray's OnResize body (main.go:89-137)
        = Render + downscale + ANSI frame on the device, only the ANSI bytes come back."""
        ansi = [0]

        def step():
            ctx.upload(fl)
            if interactive:
                st_ = ctx.render(cam, prm, None)                                         # frame stays in HBM
                frame, _, _ = ctx.present(term_cols, term_rows * 2, want_image=False)    # BiLinear s=4 + half-block ANSI on device
                ansi[0] = len(frame)
            elif with_exchange:
                st_ = ctx.render(cam, prm, None)
                exchange(img)
            else:
                st_ = ctx.render(cam, prm, img)
            return st_
        for _ in range(warm):
            step()
        sync_all()
        t0 = time.perf_counter()
        n_paths, st_ = 0, None
        for _ in range(steps):
            st_ = step()
            n_paths += st_["paths"]
        sync_all()
        ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        return sum_over_ranks(n_paths) / (ms * 1e-3) / 1e6, ms / steps, st_, ansi[0]

    def h2d_bytes(fl):
        return int(sum(fl[k].nbytes for k in ("cx", "cy", "cz", "r", "kind", "params")) + 48 +
                   __import__("ctypes").sizeof(_lib.CameraC) + __import__("ctypes").sizeof(_lib.Params))

    interactive = workload == "config5" and world == 1
    e2e_value, e2e_ms_step, st, ansi_bytes = e2e_run(cam_c, params, flat, host_img, args.steps, min(args.warmup, 2), interactive, split_samples)
    if interactive:
        ctx.read_image(host_img)  # the frame stayed in HBM (only the ANSI bytes crossed PCIe): fetch it for the parity check below
    h2d = h2d_bytes(flat)
    my_rows = st["paths"] // (w * spp)
    d2h = int(ansi_bytes) if interactive else (int(h * w * 4) if split_samples else int(my_rows * w * 4))

    # ---- the -save path: PNG of the frame encoded on the device (outside every timed region above) ----
    save_png = None
    if world == 1 and not interactive and not args.no_alt:
        import io
        ctx.render(cam_c, params, None)
        ctx.encode_png(w, h)  # warm-up (lazy kernel loading, work buffers)
        t0 = time.perf_counter()
        png, png_ms = ctx.encode_png(w, h)
        png_wall = (time.perf_counter() - t0) * 1e3
        save_png = {"bytes": len(png), "device_ms": png_ms, "wall_ms_incl_d2h": png_wall, "raw_rgb_bytes": w * h * 3,
                    "note": "tray_encode_png: adaptive filter + dynamic-Huffman deflate + Adler-32/CRC-32 on the GPU; only the file crosses PCIe"}
        try:
            from PIL import Image
            t0 = time.perf_counter()
            buf = io.BytesIO()
            Image.fromarray(host_img[:, :, :3]).save(buf, format="PNG")
            save_png["cpu_pillow_ms"] = (time.perf_counter() - t0) * 1e3
            save_png["cpu_pillow_bytes"] = buf.tell()
            save_png["decoded_equal"] = bool(np.array_equal(np.asarray(Image.open(io.BytesIO(png)).convert("RGB")), host_img[:, :, :3]))
        except Exception as e:  # Pillow missing: the device numbers stand alone
            save_png["cpu_pillow_ms"] = None
            save_png["cpu_note"] = str(e)

    threads = host_threads()
    # ---- N > 1: the other multi-GPU variants, after the main timed legs ----
    multi_gpu = {}
    if world > 1 and not args.no_alt:
        main_img = np.array(host_img) if rank == 0 else None  # the multi-process tile image (all ranks' rows are in after sync_all)
        # (c) sample split + NCCL sum-reduce, a short leg on all ranks (skipped when it IS the main mode)
        if not split_samples and spp >= world:
            ps = sample_params()
            ctx.render(cam_c, ps, None); exchange()
            sync_all()
            a = timed_steps(cam_c, ps, 2, True)
            a_ms = max_over_ranks(a["ms"])
            img_s = np.zeros((h, w, 4), dtype=np.uint8) if rank == 0 else None
            ctx.render(cam_c, ps, None); exchange(img_s)
            rec = {"precision": "fp64, samples%d+nccl_reduce" % world, "value": sum_over_ranks(a["paths"]) / (a_ms * 1e-3) / 1e6, "unit": "Mpaths/s",
                   "exchange": {"collective": "NCCL reduce(sum) to rank 0 + resolve", "bytes": int(w) * h * 24, "ms_per_step": a["exchange_ms"] / 2},
                   "note": "rank r traces samples s == r (mod N) of every pixel; ONE exchange step: fp64 colour sums (w*h*3 doubles) reduced in place "
                           "on the library's device buffer, rank 0 applies 1/N + sRGB (bench.py --split samples makes this the main mode)"}
            if rank == 0:
                dd = np.abs(img_s[:, :, :3].astype(np.int16) - main_img[:, :, :3].astype(np.int16)).max(axis=2)
                rec["vs_tiles_image"] = {"pixels_identical_frac": float((dd == 0).mean()), "pixels_within_1lsb_frac": float((dd <= 1).mean())}
                rec["not_more_than_2pct_slower_than_tiles"] = bool(rec["value"] / value >= 0.98)
            alts.append(rec)
        host_barrier()
        if rank == 0:
            # (a) config-size parity: strided rows of the multi-process image against the CPU oracle
            try:
                multi_gpu["parity"], _ = oracle_rows_check(workload, main_img, threads, threads)  # one row per host thread: ~1 Mpaths each at config 3
            except Exception as e:
                multi_gpu["parity_error"] = str(e)
            # (b) ONE process, ONE context over all N devices: what the cgo shim's TRAY_GPUS uses (Tracer.Render's fan-out,
            #     ray/tracer.go:85-116): interleaved tiles, and the sample split with combine_kernel over peer memory
            try:
                ctx_g = ray.Context(list(range(world)))
                ctx_g.upload(flat)
                inctx = {}
                for name, split in (("tiles", ray.SPLIT_TILES), ("samples", ray.SPLIT_SAMPLES)):
                    pg = tr._params(0, h)
                    pg.shard_index, pg.shard_count, pg.split_mode = 0, 0, split
                    img_g = np.zeros((h, w, 4), dtype=np.uint8)
                    ctx_g.render(cam_c, pg, None)
                    g_ms, g_paths, k = 0.0, 0, 2
                    for _ in range(k):
                        stg = ctx_g.render(cam_c, pg, None)
                        g_ms += stg["kernel_ms"]; g_paths += stg["paths"]
                    ctx_g.render(cam_c, pg, img_g)  # (untimed: the first render into a host image allocates the pinned staging of every device)
                    t0 = time.perf_counter()
                    stg = ctx_g.render(cam_c, pg, img_g)
                    g_wall = (time.perf_counter() - t0) * 1e3
                    dd = np.abs(img_g[:, :, :3].astype(np.int16) - main_img[:, :, :3].astype(np.int16)).max(axis=2)
                    v = g_paths / (g_ms * 1e-3) / 1e6
                    inctx[name] = {"value": v, "unit": "Mpaths/s", "ms_per_step": g_ms / k, "e2e_value": stg["paths"] / (g_wall * 1e-3) / 1e6,
                                   "n_devices": int(stg["n_devices"]), "pixels_identical_frac": float((dd == 0).mean()),
                                   "pixels_within_1lsb_frac": float((dd <= 1).mean()), "not_more_than_2pct_slower_than_multiprocess_tiles": bool(v / value >= 0.98)}
                inctx["note"] = ("one process drives all %d GPUs through one tray_ctx (tray_init(devices, n)): 'tiles' = interleaved rows, device-to-host "
                                 "gather only, must equal the multi-process image bit for bit; 'samples' = device g traces samples s == g (mod G), combine_kernel on "
                                 "device 0 sums the partial sums through peer pointers over NVLink and applies 1/N + sRGB (<= 1 LSB); device time = max over devices") % world
                multi_gpu["in_context"] = inctx
                ctx_g.close()
            except Exception as e:
                multi_gpu["in_context_error"] = str(e)
            # (d) the same workload on ONE GPU of this box, for a like-for-like scaling ratio inside this run
            try:
                p1 = tr._params(0, h)
                p1.shard_index, p1.shard_count = 0, 0
                ctx.render(cam_c, p1, None)
                st1 = ctx.render(cam_c, p1, None)
                v1 = st1["paths"] / (st1["kernel_ms"] * 1e-3) / 1e6
                multi_gpu["same_workload_one_gpu"] = {"value": v1, "unit": "Mpaths/s", "ms_per_step": st1["kernel_ms"], "speedup_of_this_run": value / v1,
                                                      "note": "rank 0 alone renders the whole %s frame (the driver's N=1 line is config2)" % workload}
            except Exception as e:
                multi_gpu["same_workload_one_gpu_error"] = str(e)
        host_barrier()

    # ---- CPU baseline + parity spot check (rank 0, N=1 only) ----
    cpu_baseline, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cp = cpu_port_run(workload, threads, args.cpu_seconds, fma_mode=0)
        img, cst, cdt = cpu_port_step(cp)
        sample = "rows y%%%d==0 of %s (%d rows, %.2f Mpaths), per-sample streams, strict fp64, %.1f s" % (
            cp["step"], workload, cp["rows"], cst["paths"] / 1e6, cdt)
        cpu_baseline = {"value": cst["paths"] / cdt / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port", "sample": sample}
        rows = list(range(0, h, cp["step"]))
        d = np.abs(img[rows, :, :3].astype(np.int16) - host_img[rows, :, :3].astype(np.int16)).max(axis=2)
        parity = {"rows_checked": len(rows), "pixels_identical_frac": float((d == 0).mean()), "pixels_within_1lsb_frac": float((d <= 1).mean()),
                  "mode": "GPU %s vs strict CPU port (oracle)" % args.precision}

    # ---- N = 1: every other BASELINE config, short runs, each with value, e2e and a parity spot check ----
    configs = {}
    if world == 1 and workload == "config2" and not args.no_configs and not args.no_cpu_baseline:
        for name, k_dev, k_e2e, n_rows in (("config1", 5, 5, 45), ("config3", 2, 1, threads), ("config4", 3, 2, max(1, threads // 2)), ("config5", 5, 5, 45)):
            try:
                t2, fl2 = job(name)
                w2, h2, spp2, depth2, _, desc2 = WORKLOADS[name]
                ctx.upload(fl2)
                cam2, p2 = t2.to_c(), t2._params(0, h2)
                img2 = np.zeros((h2, w2, 4), dtype=np.uint8)
                for _ in range(2):
                    ctx.render(cam2, p2, None)
                a = timed_steps(cam2, p2, k_dev)
                inter = name == "config5"
                e_val, e_ms, st2, ansi2 = e2e_run(cam2, p2, fl2, img2, k_e2e, 1, inter)
                if inter:
                    ctx.read_image(img2)
                par, cst2 = oracle_rows_check(name, img2, n_rows, threads)
                rec = {"desc": desc2, "spheres": len(fl2["cx"]), "value": a["paths"] / (a["ms"] * 1e-3) / 1e6, "unit": "Mpaths/s", "ms_per_step": a["ms"] / k_dev,
                       "mrays_per_s": a["segments"] / (a["ms"] * 1e-3) / 1e6, "steps": k_dev,
                       "sphere_tests_per_segment": a["sphere_tests"] / max(1, a["segments"]),
                       "e2e": {"value": e_val, "unit": "Mpaths/s", "ms_per_step": e_ms, "h2d_bytes_per_step": h2d_bytes(fl2),
                               "d2h_bytes_per_step": int(ansi2) if inter else int(h2 * w2 * 4)},
                       "parity_spot_check": par,
                       "cpu_port": {"value": cst2["paths"] / cst2["seconds"] / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port"}}
                if inter:
                    rec["keypress_latency_ms"] = e_ms
                    rec["interactive"] = {"terminal": "%dx%d" % (term_cols, term_rows), "supersample": 4, "ansi_frame_bytes": int(ansi2)}
                configs[name] = rec
            except Exception as e:
                configs[name] = {"error": str(e)}
        ctx.upload(flat)

    if rank == 0:
        traffic, traffic_src = ncu_traffic_bytes() if workload == "config2" and args.precision == "fp64" else (None, None)
        prof = ncu_summary() if workload == "config2" and args.precision == "fp64" else None
        ns_seg = trace_ms * 1e6 / max(1, segments) * (148 * 4)  # SM-sub-partition nanoseconds per ray segment
        line = {
            "metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "fp32" else "f64", "data": "synthetic",
            "mrays_per_s": all_segments / (dev_ms * 1e-3) / 1e6,
            "ns_per_segment": trace_ms * 1e6 / max(1, segments),
            "wall_ms_per_step": wall_ms / args.steps,
            "config": {"workload": workload, "desc": desc, "width": w, "height": h, "rays_per_pixel": spp, "max_depth": depth,
                       "seed": SEED, "spheres": n_spheres, "precision": args.precision, "streams": "per-sample", "layout": args.layout,
                       "closest_hit": structure,
                       "parallelism": ("samples%d+nccl_reduce" if split_samples else "tiles%d") % world if world > 1 else "1gpu",
                       "l2": "256 MiB memset between timed steps (L2 flush); working set is 12 kB of sphere tables in shared memory"},
            "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                       "power_w_max": clocks.get("power_w_max"), "samples": clocks["samples"]},
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": peak_used, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_used if peak_used else None, "traffic": traffic,
                         "traffic_note": "DRAM bytes per trace_kernel launch from %s (ncu --set full); algorithmic DRAM bytes per path: 24 B of "
                                         "sample colour written (read back once by the resolve kernel)" % traffic_src if traffic else None,
                         "kernel": "tray::trace_kernel", "launches": int(trace_launches), "avg_launch_ms": trace_ms / max(1, trace_launches),
                         "algorithmic_flops_per_launch": flops / max(1, trace_launches),
                         "flops_model": "SURVEY 8(d): 18 flops per sphere test + 155 per segment, N_tested = kernel-counted tests for a hierarchy (N for the "
                                        "linear scan): %.1f of %d spheres (and %.1f boxes, not counted) looked at per segment (%s). The tests run in packed "
                                        "fp32 behind an exact filter, so this is not FP64-pipe utilisation either: see fp64_pipe_busy, issue_active, "
                                        "frac_fp64_brute (the all-fp64 linear scan, the kernel that sits on the FP64 roofline) and ns_per_segment; "
                                        "frac_reference_equivalent_work counts the reference algorithm's N tests per Scene.Hit instead" % (
                                            sphere_tests / max(1, segments), n_spheres, box_tests / max(1, segments), structure),
                         "peak_source": "measured live on this GPU by tray_measure_peak (DFMA chains, 8/thread); MEASURED_PEAKS.json has no fp64 entry",
                         "peak_dadd_dmul_tflops": peak_strict_tf, "peak_ffma_tflops": peak_f32_tf, "peak_ffma2_tflops": peak_ffma2_tf,
                         "loop_only_probe_tflops": loop_probe_tf,
                         "structural_ceiling_frac": {"fp64": 18.0 / 34.0, "fp64-brute": 18.0 / 34.0, "fp64-fma": 18.0 / 22.0, "fp32": 18.0 / 22.0}[args.precision],
                         # what the hardware is actually doing (committed ncu capture of this command line) -- VERDICT r1 item 4
                         "fp64_pipe_busy": prof.get("fp64_pipe_busy") if prof else None,
                         "fp32_pipe_slot_frac": prof.get("fp32_pipe_slot_frac") if prof else None,
                         "issue_active": prof.get("issue_active") if prof else None,
                         "threads_per_instruction": prof.get("threads_per_instruction") if prof else None,
                         "frac_fp64_brute": frac_fp64_brute,
                         # the same formula with the reference algorithm's N tests per Scene.Hit: how fast reference-equivalent work is retired
                         "frac_reference_equivalent_work": (flops_reference / (trace_ms * 1e-3) / 1e12 / peak_used if peak_used and trace_ms > 0 else None),
                         "ns_per_segment": ({"total_smsp_ns": ns_seg, **{k: ns_seg * v for k, v in prof.get("time_share", {}).items()},
                                             "note": "SM-sub-partition time per ray segment (trace-kernel time x 592 sub-partitions / segments), split by the "
                                                     "stall-sample shares of the committed capture (%s)" % prof.get("source")} if prof else {"total_smsp_ns": ns_seg}),
                         "note": "compute/latency-bound: HBM traffic is ~30 B/path (sample colour out); tensor cores do not apply"},
            "segments_per_path": all_segments / all_paths,
        }
        if per_rank:
            line["per_rank"] = per_rank
        if split_samples:
            line["exchange"] = {"collective": "NCCL reduce(sum) to rank 0 + resolve", "bytes": int(w) * h * 24, "ms_per_step": exchange_ms / args.steps,
                                "note": "fp64 colour sums (w*h*3 doubles) reduced in place on the library's device buffer; rank 0's time"}
        if interactive:
            line["keypress_latency_ms"] = e2e_ms_step
            line["interactive"] = {"terminal": "%dx%d" % (term_cols, term_rows), "supersample": 4, "ansi_frame_bytes": int(ansi_bytes),
                                   "note": "per-keypress OnResize body: scene upload + Render + on-device BiLinear downscale + half-block ANSI "
                                           "frame, only the ANSI bytes cross PCIe"}
        if alts:
            line["alt_modes"] = alts
        if save_png:
            line["save_png"] = save_png
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if parity:
            line["parity_spot_check"] = parity
        if multi_gpu.get("parity"):
            line["parity_spot_check"] = multi_gpu.pop("parity")
        for k2, v2 in multi_gpu.items():
            line[k2] = v2
        if configs:
            line["configs"] = configs
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if rank == 0:
            try:
                os.unlink(shm)
            except OSError:
                pass
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp64-brute", "fp64-fma", "fp32"],
                    help="fp64 = strict Go/amd64 semantics with the exact fp32 pre-filter (default); fp64-brute = same, all tests in fp64; "
                         "fp64-fma = fused discriminant; fp32 = fast path")
    ap.add_argument("--split", default="tiles", choices=["tiles", "samples"],
                    help="N>1 partitioning: interleaved row bands (no collective, default) or sample split + NCCL sum-reduce")
    ap.add_argument("--layout", default="auto", choices=["auto", "plain", "regroup", "wavefront"],
                    help="divergence layout of the trace kernel (results identical): plain megakernel or per-material regrouping")
    ap.add_argument("--no-alt", action="store_true", help="skip the alternative kernels / layouts / multi-GPU variants")
    ap.add_argument("--no-configs", action="store_true", help="N=1: skip the short runs of the other BASELINE configs")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="--impl reference: seconds of CPU work for the whole run (bounded sample per step)")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import io
    buf = io.StringIO()
    sys.stdout = buf
    try:
        rc = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout = sys.__stdout__
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    out = buf.getvalue()
    if out:
        sys.stdout.write(out)
        sys.stdout.flush()
    return rc


if __name__ == "__main__":
    sys.exit(main())

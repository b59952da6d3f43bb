"""Ad-hoc first GPU check (superseded by tests/ -m gpu): parity probes + timing."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tray_b200 import ray, rand
from oracle import oracle as O

ctx = ray.default_context()
print("peak dfma/dadd+dmul/ffma TF:", [ctx.measure_peak(k) for k in (0, 1, 2)])
# rng
for kind, fn in ((0, O.rng_u64), (1, O.rng_f64), (2, O.rng_norm)):
    a = ctx.rng_dump(kind, 5, 42, 20000); b = fn(5, 42, 20000)
    print("rng kind", kind, "equal:", np.array_equal(a.view(np.uint64), b.view(np.uint64)))
a = ctx.rng_dump(3, 7, 42, 5000); b = O.rng_unit_vectors(7, 42, 5000); print("unitvec equal", np.array_equal(a, b))
a = ctx.rng_dump(4, 7, 42, 5000, 0.5); b = O.rng_in_disc(7, 42, 0.5, 5000); print("indisc equal", np.array_equal(a, b))
x = np.concatenate([np.linspace(-0.1, 1.1, 100001), np.random.default_rng(1).random(200000)])
print("srgb equal", np.array_equal(ctx.linear_to_srgb(x), np.array([O.linear_to_srgb(v) for v in x], dtype=np.uint8)))

seed = 2
scene = ray.RichScene(rand.New(seed)); osc = O.rich_scene(seed)
def tracer(w, h, spp, d, prec, mode=ray.STREAM_PER_SAMPLE, workers=0):
    t = ray.New(w, h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed = d, spp, seed
    t.Precision, t.StreamMode, t.NumWorkers = prec, mode, workers
    return t
# first hit
w, h = 400, 225
t = tracer(w, h, 1, 1, ray.FP64_STRICT); t.Initialize(w, h); ctx.upload(scene.flatten())
cam = O.camera_init(w, h, **O.RICH_CAMERA)
for prec, fma in ((ray.FP64_STRICT, 0), (ray.FP64_FMA, 1)):
    g = ctx.first_hit(t.to_c(), w, h, prec); o = O.first_hit(osc, cam, w, h, fma)
    print("first_hit prec", prec, [np.array_equal(x, y) for x, y in zip(g, o)])
# render parity config 1
for prec, fma in ((ray.FP64_FMA, 1), (ray.FP64_STRICT, 0)):
    t = tracer(w, h, 10, 50, prec); t0 = time.time(); img = t.Render(scene).copy(); dt = time.time() - t0
    p = O.make_params(w, h, spp=10, max_depth=50, seed=seed, num_workers=8, stream_mode=1, fma_mode=fma)
    ref, hdr, st = O.render(osc, cam, p, want_hdr=True)
    ghdr = ctx.read_hdr(w, h)
    print("render prec", prec, "img equal", np.array_equal(img, ref), "hdr equal", np.array_equal(ghdr, hdr),
          "ndiff", int((img != ref).any(axis=2).sum()), "seg", t.Stats["segments"], st["segments"], "wall", dt, t.Stats)
# conformance: reference streams, 8 workers and 1 worker (small)
for workers, (cw, ch, spp) in ((8, (400, 225, 10)), (1, (120, 68, 4))):
    t = tracer(cw, ch, spp, 50, ray.FP64_STRICT, ray.STREAM_REFERENCE, workers); t0 = time.time(); img = t.Render(scene).copy(); dt = time.time() - t0
    c2 = O.camera_init(cw, ch, **O.RICH_CAMERA)
    p = O.make_params(cw, ch, spp=spp, max_depth=50, seed=seed, num_workers=workers, stream_mode=0, fma_mode=0)
    ref, _, st = O.render(osc, c2, p)
    print("conformance workers", workers, "equal", np.array_equal(img, ref), "ndiff", int((img != ref).any(axis=2).sum()), "wall", dt, t.Stats["kernel_ms"])
# timing config 2
for prec in (ray.FP64_FMA, ray.FP64_STRICT, ray.FP32):
    t = tracer(1920, 1080, 64, 50, prec)
    for rep in range(2):
        t0 = time.time(); t.Render(scene); dt = time.time() - t0
    s = t.Stats
    flops = s["segments"] * (18.0 * 485 + 155)
    print("config2 prec", prec, "wall %.3f s kernel %.1f ms trace %.1f ms" % (dt, s["kernel_ms"], s["trace_kernel_ms"]),
          "Mpaths/s %.1f Mrays/s %.1f TFLOP/s(alg) %.2f" % (s["paths"] / s["kernel_ms"] / 1e3, s["segments"] / s["kernel_ms"] / 1e3, flops / s["kernel_ms"] / 1e9), "launches", s["launches"])

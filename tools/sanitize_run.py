"""Small invocation of every device path, meant to run under compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py

Covers the race-prone parts named in VERDICT r1: the regroup exchange area, the wavefront queues, the shared-memory camera-ray
pool and parking area of the megakernel, the cluster tables (every chunk-count remainder, so that the scan's table reads are
shown to stay inside the dynamic allocation for every padded size), the device LBVH build, present and the PNG encoder."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tray_b200 import _lib, rand, ray  # noqa: E402

w, h, spp, depth = 48, 27, 3, 8
scene = ray.RichScene(rand.New(2))
ctx = ray.default_context()
imgs = {}
for prec in (ray.FP64_STRICT, ray.FP64_STRICT_BRUTE, ray.FP64_FMA, ray.FP32):
    for layout in (ray.LAYOUT_PLAIN, ray.LAYOUT_REGROUP, ray.LAYOUT_WAVEFRONT):
        for accel in (ray.ACCEL_AUTO, ray.ACCEL_BRUTE, ray.ACCEL_BVH, ray.ACCEL_CLUSTER):
            t = ray.New(w, h)
            t.Camera = ray.RichSceneCamera()
            t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision, t.Layout, t.Accel = depth, spp, 2, prec, layout, accel
            imgs[(prec, layout, accel)] = t.Render(scene).copy()
base = imgs[(ray.FP64_STRICT, ray.LAYOUT_PLAIN, ray.ACCEL_BRUTE)]
for k, v in imgs.items():
    if k[0] in (ray.FP64_STRICT, ray.FP64_STRICT_BRUTE):
        assert np.array_equal(v, base), k
# every table size n = 0..20 and around the group boundaries: padded chunks, partial groups, always-groups only
rs = np.random.RandomState(3)
for n in list(range(0, 21)) + [63, 64, 65, 127, 129, 511, 513]:
    objs = [ray.Sphere(tuple(rs.uniform(-4, 4, 3) + np.array([0, 0, -8.0])), float(rs.uniform(0.1, 0.8)),
                       (ray.Lambertian((.5, .6, .7)), ray.Metal((.8, .8, .8), 0.2), ray.Dielectric(1.5))[i % 3]) for i in range(n)]
    if n % 4 == 1:
        objs.append(ray.Sphere((0, -1000.5, -8), 1000, ray.Lambertian((.5, .5, .5))))  # outside the filter's range: always-group
    sc = ray.Scene(objs, ray.DefaultBackground())
    out = {}
    for accel in (ray.ACCEL_CLUSTER, ray.ACCEL_BRUTE):
        t = ray.New(40, 22)
        t.Camera = ray.Camera(VerticalFoV=50.0)
        t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Accel = 6, 2, 5, accel
        out[accel] = t.Render(sc).copy()
    assert np.array_equal(out[ray.ACCEL_CLUSTER], out[ray.ACCEL_BRUTE]), n
# scenes above 2048 spheres: the three-level walk with its tables in global memory (cluster_scan_big), word counts 5, 6 and 9 (one and two
# octets of word boxes, partial last word), with an always-group
for n in (2049, 2600, 4103):
    objs = [ray.Sphere((float(rs.uniform(-30, 30)), float(rs.uniform(0, 2)), float(rs.uniform(-60, 0))), float(rs.uniform(0.1, 0.5)),
                       (ray.Lambertian((.5, .6, .7)), ray.Metal((.8, .8, .8), 0.2), ray.Dielectric(1.5))[i % 3]) for i in range(n)]
    objs.append(ray.Sphere((0, -1000.5, -8), 1000, ray.Lambertian((.5, .5, .5))))
    sc = ray.Scene(objs, ray.DefaultBackground())
    out = {}
    for accel in (ray.ACCEL_AUTO, ray.ACCEL_CLUSTER, ray.ACCEL_BVH, ray.ACCEL_BRUTE):
        t = ray.New(40, 22)
        t.Camera = ray.Camera(Position=(0, 3, 6), LookAt=(0, 0, -20), VerticalFoV=50.0)
        t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Accel = 6, 2, 5, accel
        out[accel] = t.Render(sc).copy()
        if accel in (ray.ACCEL_AUTO, ray.ACCEL_CLUSTER):
            assert t.Stats["box_tests"] > 0, n
    assert all(np.array_equal(v, out[ray.ACCEL_BRUTE]) for v in out.values()), n
# device LBVH build + traversal on a scene large enough for it
big = ray.RichScene(rand.New(2), 20)
ctx.configure(_lib.CFG_BVH_BUILD, _lib.BVH_BUILD_DEVICE)
t = ray.New(32, 18)
t.Camera = ray.RichSceneCamera()
t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Accel = 6, 2, 2, ray.ACCEL_BVH
t.Render(big)
ctx.configure(_lib.CFG_BVH_BUILD, _lib.BVH_BUILD_AUTO)
t = ray.New(w, h)
t.Camera = ray.RichSceneCamera()
t.MaxDepth, t.NumRaysPerPixel, t.Seed = depth, spp, 2
for done, img in t.RenderProgressive(scene, 2):
    pass
assert np.array_equal(img, base)
t.Render(scene)
ctx.present(16, 10)
ctx.present(96, 54)
png, _ = ctx.encode_png(w, h)
t.StreamMode, t.NumWorkers = ray.STREAM_REFERENCE, 3
t.Render(scene)
t.StreamMode, t.MaxDepth = ray.STREAM_PER_SAMPLE, 0   # MaxDepth 0: every pixel (0,0,0,255)
p = t._params(0, h)
p.max_depth = 0
ctx.render(t.to_c(), p, t.imageData)
assert (t.imageData[:, :, :3] == 0).all() and (t.imageData[:, :, 3] == 255).all()
ctx.first_hit(t.to_c(), w, h)
for v in range(3):
    ctx.configure(_lib.CFG_INDISC, v); ctx.configure(_lib.CFG_UNITVEC, v)
    ctx.rng_dump(3, 1, 2, 16); ctx.rng_dump(4, 1, 2, 16)
ctx.configure(_lib.CFG_INDISC, 0); ctx.configure(_lib.CFG_UNITVEC, 0)
print("sanitize_run ok", len(png))

"""Small invocation of every device path, meant to run under compute-sanitizer (memcheck / racecheck / initcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from tray_b200 import rand, ray  # noqa: E402

w, h, spp, depth = 48, 27, 3, 8
scene = ray.RichScene(rand.New(2))
ctx = ray.default_context()
imgs = {}
for prec in (ray.FP64_STRICT, ray.FP64_STRICT_BRUTE, ray.FP64_FMA, ray.FP32):
    for layout in (ray.LAYOUT_PLAIN, ray.LAYOUT_REGROUP):
        for accel in (ray.ACCEL_BRUTE, ray.ACCEL_BVH):
            t = ray.New(w, h)
            t.Camera = ray.RichSceneCamera()
            t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision, t.Layout, t.Accel = depth, spp, 2, prec, layout, accel
            imgs[(prec, layout, accel)] = t.Render(scene).copy()
base = imgs[(ray.FP64_STRICT, ray.LAYOUT_PLAIN, ray.ACCEL_BRUTE)]
for k, v in imgs.items():
    if k[0] in (ray.FP64_STRICT, ray.FP64_STRICT_BRUTE):
        assert np.array_equal(v, base), k
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
ctx.first_hit(t.to_c(), w, h)
ctx.rng_dump(3, 1, 2, 16)
print("sanitize_run ok", len(png))

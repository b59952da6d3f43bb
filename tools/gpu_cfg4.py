"""Tuning aid: closest-hit structures on BASELINE config 2 and config 4 (10 001 spheres), whole frame, image hash."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_b200 import ray, rand
for half, w, h, spp, d in ((11, 1920, 1080, 64, 50), (50, 1920, 1080, 64, 12)):
    scene = ray.RichScene(rand.New(2), half)
    for accel, name in ((ray.ACCEL_AUTO, "auto"), (ray.ACCEL_CLUSTER, "cluster"), (ray.ACCEL_BVH, "bvh"), (ray.ACCEL_BRUTE, "brute")):
        spp_use = 4 if (half == 50 and accel == ray.ACCEL_BRUTE) else spp
        t = ray.New(w, h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Accel = d, spp_use, 2, accel
        best = 1e9
        for rep in range(3):
            t.Render(scene); best = min(best, t.Stats["kernel_ms"])
        s = t.Stats
        print("n=%d %-7s spp %d: %.1f ms, %.1f Mpaths/s, %.1f Mrays/s, sphere tests/segment %.1f, boxes/segment %.1f, sha %s" % (
            len(scene.Objects), name, spp_use, best, s["paths"] / best / 1e3, s["segments"] / best / 1e3,
            s["sphere_tests"] / s["segments"], s["box_tests"] / s["segments"], hashlib.sha1(t.imageData.tobytes()).hexdigest()[:10]), flush=True)

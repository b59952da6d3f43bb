import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tray_b200 import ray, rand
for half, w, h, spp, d in ((11, 1920, 1080, 64, 50), (50, 1920, 1080, 64, 12)):
    scene = ray.RichScene(rand.New(2), half)
    for accel, name in ((ray.ACCEL_BVH, "bvh"), (ray.ACCEL_BRUTE, "brute")):
        if half == 50 and accel == ray.ACCEL_BRUTE: spp_use = 4
        else: spp_use = spp
        t = ray.New(w, h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Accel = d, spp_use, 2, accel
        for rep in range(2): t.Render(scene)
        s = t.Stats
        print("n=%d %s spp %d: %.1f ms, %.1f Mpaths/s, %.1f Mrays/s, tests/segment %.1f" % (len(scene.Objects), name, spp_use, s["kernel_ms"], s["paths"]/s["kernel_ms"]/1e3, s["segments"]/s["kernel_ms"]/1e3, s["sphere_tests"]/s["segments"]))

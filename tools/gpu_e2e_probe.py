"""Tuning aid: where the end-to-end time of a config-2 frame goes beyond the kernels (upload, render into a host buffer vs device-resident)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tray_b200 import ray, rand
scene = ray.RichScene(rand.New(2))
t = ray.New(1920, 1080); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed = 50, 64, 2
t.Render(scene)
ctx = ray.default_context()
flat = scene.flatten()
cam = t.to_c(); prm = t._params(0, 1080)
img = np.zeros((1080, 1920, 4), dtype=np.uint8)
def timeit(f, n=20):
    f(); t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e3
print("upload only            %.3f ms" % timeit(lambda: ctx.upload(flat)))
st = {}
def r_dev(): st.update(ctx.render(cam, prm, None))
def r_host(): st.update(ctx.render(cam, prm, img))
a = timeit(r_dev, 10); print("render, image left on device  %.3f ms wall, kernels %.3f" % (a, st["kernel_ms"]))
b = timeit(r_host, 10); print("render into host buffer       %.3f ms wall, kernels %.3f" % (b, st["kernel_ms"]))
def both(): ctx.upload(flat); r_host()
c = timeit(both, 10); print("upload + render into host     %.3f ms wall" % c)

"""Tuning aid: the fp32 fast path of every library variant in build/variants/ (and of the in-tree library) on config 2:
device time, Mpaths/s, PSNR of its 8-bit image against the fp64 image of the same library, and the fp64 default beside it
(time + image hash: the default kernel must not move)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, hashlib
sys.path.insert(0, %r)
import numpy as np
from tray_b200 import ray, rand
scene = ray.RichScene(rand.New(2))
out = {}
imgs = {}
for name, prec in (("fp64", ray.FP64_STRICT), ("fp32", ray.FP32)):
    t = ray.New(1920, 1080); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision = 50, 64, 2, prec
    best = 1e9
    for rep in range(3):
        t.Render(scene); best = min(best, t.Stats["trace_kernel_ms"])
    imgs[name] = t.imageData.copy()
    out[name] = dict(ms=round(best, 3), mpaths=round(t.Stats["paths"] / best / 1e3, 1), seg=round(t.Stats["segments"] / t.Stats["paths"], 4),
                     sha=hashlib.sha1(t.imageData.tobytes()).hexdigest()[:8])
a = imgs["fp64"][..., :3].astype(np.float64); b = imgs["fp32"][..., :3].astype(np.float64)
mse = float(np.mean((a - b) ** 2))
out["psnr_db"] = round(10 * np.log10(255.0 ** 2 / mse), 2) if mse > 0 else None
out["within_1lsb"] = round(float(np.mean(np.abs(a - b) <= 1)), 4)
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "tray_b200", "libtraycuda.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, TRAY_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)

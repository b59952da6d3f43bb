"""Tuning aid: whole-frame time (all kernels) and image hash of every library variant in build/variants/ on config 2,
for the plain and the regroup layout."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, hashlib
sys.path.insert(0, %r)
from tray_b200 import ray, rand
scene = ray.RichScene(rand.New(2))
out = {}
for name, layout in (("plain", ray.LAYOUT_PLAIN), ("regroup", ray.LAYOUT_REGROUP)):
    t = ray.New(1920, 1080); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed = 50, 64, 2
    t.Layout = layout
    best = 1e9
    for rep in range(4):
        t.Render(scene); best = min(best, t.Stats["kernel_ms"])
    out[name] = dict(ms=round(best, 2), mpaths=round(t.Stats["paths"] / best / 1e3, 1), sha=hashlib.sha1(t.imageData.tobytes()).hexdigest()[:10])
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "tray_b200", "libtraycuda.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, TRAY_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)

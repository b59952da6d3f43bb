"""Tuning aid: whole-frame time (all kernels) and image hash of every library variant in build/variants/ on config 2."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
from tray_b200 import ray, rand
scene = ray.RichScene(rand.New(2))
t = ray.New(1920, 1080); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed = 50, 64, 2
best = 1e9
for rep in range(3):
    t.Render(scene); best = min(best, t.Stats["kernel_ms"])
import hashlib
print(json.dumps(dict(ms=best, mpaths=t.Stats["paths"] / best / 1e3, launches=t.Stats["launches"], segments=t.Stats["segments"],
                      image_sha=hashlib.sha1(t.imageData.tobytes()).hexdigest()[:12])))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "tray_b200", "libtraycuda.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, TRAY_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])

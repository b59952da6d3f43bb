"""Tuning aid: device time of the small BASELINE frames (config 1: 400x225x10, config 5: 640x360x64 depth 12) and config 2 for every library variant."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, hashlib
sys.path.insert(0, %r)
from tray_b200 import ray, rand
scene = ray.RichScene(rand.New(2))
out = {}
for name, (w, h, spp, d) in (("config1", (400, 225, 10, 50)), ("config5", (640, 360, 64, 12)), ("config2", (1920, 1080, 64, 50))):
    t = ray.New(w, h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed = d, spp, 2
    best = 1e9
    for rep in range(8 if w < 1000 else 3):
        t.Render(scene); best = min(best, t.Stats["kernel_ms"])
    out[name] = dict(ms=round(best, 3), mpaths=round(t.Stats["paths"] / best / 1e3, 1), sha=hashlib.sha1(t.imageData.tobytes()).hexdigest()[:8])
print(json.dumps(out))
''' % ROOT
for lib in sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "tray_b200", "libtraycuda.so")]:
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, TRAY_LIB=lib), capture_output=True, text=True)
    print(os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)

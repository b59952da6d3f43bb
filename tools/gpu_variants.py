"""Tuning aid: time each libtraycuda variant in build/variants/ on config 2 (and check config 1 parity)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json, time
sys.path.insert(0, %r)
import numpy as np
from tray_b200 import ray, rand
from oracle import oracle as O
scene = ray.RichScene(rand.New(2))
def tr(w,h,spp,d,prec):
    t = ray.New(w,h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision = d, spp, 2, prec
    return t
out = {}
for prec, fma, name in ((ray.FP64_STRICT_BRUTE,0,"fma"),(ray.FP64_STRICT,0,"strict")):
    t = tr(400,225,10,50,prec); img = t.Render(scene).copy()
    ref,_,st = O.render(O.rich_scene(2), O.camera_init(400,225,**O.RICH_CAMERA), O.make_params(400,225,spp=10,max_depth=50,seed=2,num_workers=8,stream_mode=1,fma_mode=fma))
    ok = bool(np.array_equal(img, ref))
    t = tr(1920,1080,64,50,prec)
    best = 1e9
    for rep in range(3):
        t.Render(scene); best = min(best, t.Stats["trace_kernel_ms"])
    s = t.Stats
    out[name] = dict(parity=ok, ms=best, mpaths=s["paths"]/best/1e3, tflops=s["segments"]*(18.0*485+155)/best/1e9)
ctx = ray.default_context()
out["probe"] = dict(fma=ctx.measure_peak(3)[0], strict=ctx.measure_peak(4)[0], dfma=ctx.measure_peak(0)[0])
print(json.dumps(out))
''' % ROOT
libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "tray_b200", "libtraycuda.so")]
for lib in libs:
    env = dict(os.environ, TRAY_LIB=lib)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
    try:
        d = json.loads(line)
        print("%-34s brute: %6.1f ms %6.1f Mp/s %5.2f TF parity=%s | filtered: %6.1f ms %6.1f Mp/s %5.2f TF parity=%s" % (
            os.path.basename(lib), d["fma"]["ms"], d["fma"]["mpaths"], d["fma"]["tflops"], d["fma"]["parity"],
            d["strict"]["ms"], d["strict"]["mpaths"], d["strict"]["tflops"], d["strict"]["parity"]),
            "| loop-only probe TF fma %.2f strict %.2f (dfma peak %.2f)" % (d["probe"]["fma"], d["probe"]["strict"], d["probe"]["dfma"]))
    except Exception:
        print(os.path.basename(lib), "FAILED", line)

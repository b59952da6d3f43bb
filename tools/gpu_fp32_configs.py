"""The fp32 fast path beside the strict fp64 default on every BASELINE config (one B200): device ms, Mpaths/s, ray segments per path,
PSNR of the 8-bit image against the fp64 image, share of pixels within 1 LSB."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from tray_b200 import ray, rand
CONFIGS = (("config1", 11, 400, 225, 10, 50), ("config2", 11, 1920, 1080, 64, 50), ("config3", 11, 3840, 2160, 256, 50),
           ("config4", 50, 1920, 1080, 64, 12), ("config5", 11, 640, 360, 64, 12))
for name, half, w, h, spp, depth in CONFIGS:
    scene = ray.RichScene(rand.New(2), half=half)
    out, imgs = {}, {}
    for label, prec in (("fp64", ray.FP64_STRICT), ("fp32", ray.FP32)):
        t = ray.New(w, h); t.Camera = ray.RichSceneCamera(); t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision = depth, spp, 2, prec
        best = 1e9
        for rep in range(2 if w > 3000 else 4):
            t.Render(scene); best = min(best, t.Stats["kernel_ms"])
        imgs[label] = t.imageData[..., :3].astype(np.float64)
        out[label] = dict(ms=round(best, 3), mpaths=round(t.Stats["paths"] / best / 1e3, 1), seg=round(t.Stats["segments"] / t.Stats["paths"], 4))
    mse = float(np.mean((imgs["fp64"] - imgs["fp32"]) ** 2))
    out["speedup"] = round(out["fp64"]["ms"] / out["fp32"]["ms"], 3)
    out["psnr_db"] = round(10 * np.log10(255.0 ** 2 / mse), 2) if mse > 0 else None
    out["within_1lsb"] = round(float(np.mean(np.abs(imgs["fp64"] - imgs["fp32"]).max(axis=2) <= 1)), 4)
    print(name, len(scene.Objects), "spheres", json.dumps(out), flush=True)

#!/usr/bin/env python3
"""Generates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists; the tests themselves never read /root/reference).

  example_png_blocks.npz   16x16 block means of the reference's only golden artefact, example.png
                           (tray -r 64 -s 8 -d 50 -seed 2, README.md:30-31): pins scene generation
                           (PCG + Float64 + seeding + RichScene draw order), the camera and the sRGB store.
  oracle_small.npz         oracle outputs on a small case (regression pin for the oracle itself and a
                           second, oracle-free target for the CUDA path).
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

ref_png = "/root/reference/example.png"
if os.path.exists(ref_png):
    from PIL import Image
    im = np.array(Image.open(ref_png))[:, :, :3].astype(np.float64)
    h, w, _ = im.shape
    blocks = im.reshape(h // 16, 16, w // 16, 16, 3).mean(axis=(1, 3))
    np.savez_compressed(os.path.join(out, "example_png_blocks.npz"), blocks=np.round(blocks * 16).astype(np.uint16),
                        width=w, height=h, note="block means x16, 16x16 px blocks of /root/reference/example.png")

W, H, SPP, D, SEED = 96, 54, 4, 50, 2
sc = O.rich_scene(SEED)
cam = O.camera_init(W, H, **O.RICH_CAMERA)
g = {}
for name, mode, fma, workers in (("ps_strict", 1, 0, 4), ("ps_fma", 1, 1, 4), ("ref_w1", 0, 0, 1), ("ref_w4", 0, 0, 4)):
    p = O.make_params(W, H, spp=SPP, max_depth=D, seed=SEED, num_workers=workers, stream_mode=mode, fma_mode=fma)
    img, hdr, st = O.render(sc, cam, p, want_hdr=True)
    g[name + "_rgba"] = img
    g[name + "_hdr"] = hdr
    g[name + "_segments"] = st["segments"]
ids, t, n, f = O.first_hit(sc, cam, W, H, 0)
g.update(fh_id=ids, fh_t=t, fh_n=n, fh_front=f)
g.update(rng_u64=O.rng_u64(5, 42, 32), rng_f64=O.rng_f64(5, 42, 32), rng_norm=O.rng_norm(5, 42, 4096),
         rng_unit=O.rng_unit_vectors(7, 42, 64), rng_disc=O.rng_in_disc(7, 42, 0.5, 64))
g.update(width=W, height=H, spp=SPP, depth=D, seed=SEED)
np.savez_compressed(os.path.join(out, "oracle_small.npz"), **g)
print("wrote", os.listdir(out))

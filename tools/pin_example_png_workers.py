#!/usr/bin/env python3
"""Extension of tools/pin_example_png.py: example.png may predate the work-queue fan-out (ray/tracer.go:90-115), i.e. come
from a static row partition with stream idx = WORKER index. For every row 300..719 as a potential block start and every
idx 0..15 (plus the seeding / InDisc / UnitVector variants), render the first pixels of the row in reference-stream mode and
compare with the PNG. Build container only (reads /root/reference/example.png). Prints candidates; none => still unpinned."""
import os, sys
import multiprocessing as mp
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

W, H, NX = 1280, 720, 6


def work(args):
    sv, indisc, uv = args
    from oracle import oracle as O
    from PIL import Image
    ref = np.array(Image.open("/root/reference/example.png"))[:, :, :3].astype(int)
    sc = O.rich_scene(2)
    cam = O.camera_init(W, H, **O.RICH_CAMERA)
    L = O.lib()
    L.oracle_set_variants(indisc, uv)
    L.oracle_set_experiment(sv, NX)
    p = O.make_params(W, H, spp=64, max_depth=50, seed=2, num_workers=11, stream_mode=0)
    hits = []
    rows = list(range(300, 720))
    for idx in range(0, 16):
        img = np.zeros((H, W, 4), dtype=np.uint8)
        for y0 in rows:
            O.render_lines(sc, cam, p, idx, y0, y0 + 1, img)
        eq = (img[rows, :NX, :3].astype(int) == ref[rows, :NX]).reshape(len(rows), -1).sum(axis=1)
        for i, y0 in enumerate(rows):
            if eq[i] >= 3 * NX - 3:  # chance level is ~25 % per channel: 15 of 18 equal does not happen by accident
                hits.append((idx, y0, int(eq[i])))
    return (sv, indisc, uv, hits)


if __name__ == "__main__":
    jobs = [(sv, i, u) for sv in range(5) for i in range(3) for u in range(3)]
    with mp.Pool(8) as pool:
        for sv, i, u, hits in pool.imap_unordered(work, jobs):
            print("seeding %d indisc %d unitvec %d: (idx,row) matches %s" % (sv, i, u, hits), flush=True)

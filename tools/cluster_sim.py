#!/usr/bin/env python3
"""Design probe for the two-level closest hit: how many 8-sphere clusters does a ray of the benchmark
scene touch, per lane and as the union over the 32 lanes of a warp? A plain numpy path tracer (not bit-exact, no parity
role) follows 32 samples of one pixel per "warp" in lockstep, like the megakernel does right after regeneration, and tests
every ray against the cluster boxes (slab test) and bounding spheres.

    python tools/cluster_sim.py [n_warps] [leaf]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O  # noqa: E402


def median_split(ids, c, leaf):
    """Chunks of <= leaf spheres, full except possibly one: split at a multiple of `leaf` nearest the median."""
    if len(ids) <= leaf:
        return [ids]
    ext = c[ids].max(axis=0) - c[ids].min(axis=0)
    ax = int(np.argmax(ext))
    order = ids[np.argsort(c[ids, ax], kind="stable")]
    n_chunks = -(-len(ids) // leaf)
    left = (n_chunks // 2) * leaf
    return median_split(order[:left], c, leaf) + median_split(order[left:], c, leaf)


def main():
    n_warps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    leaf = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(1)
    sc = O.rich_scene(2)
    c = np.stack([sc.cx, sc.cy, sc.cz], axis=1)
    r = np.asarray(sc.r)
    kind = np.asarray(sc.kind)
    prm = np.asarray(sc.params).reshape(-1, 4)
    n = len(r)
    med = np.median(r)
    big = np.where(r > 3 * med)[0]
    small = np.where(r <= 3 * med)[0]
    clusters = [big] + median_split(small, c, leaf)
    lo = np.array([(c[k] - r[k, None]).min(axis=0) for k in clusters])
    hi = np.array([(c[k] + r[k, None]).max(axis=0) for k in clusters])
    bc = 0.5 * (lo + hi)
    br = np.array([np.max(np.linalg.norm(c[k] - bc[j], axis=1) + r[k]) for j, k in enumerate(clusters)])
    print("spheres", n, "clusters", len(clusters), "leaf", leaf, "big", len(big))

    w, h = 1920, 1080
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    pos = np.array(cam.position); p00 = np.array(cam.pixel00); px = np.array(cam.pixel_x); py = np.array(cam.pixel_y)
    du = np.array(cam.defocus_u); dv = np.array(cam.defocus_v)

    def box_hits(Oo, D):
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / D
            t1 = (lo[None] - Oo[:, None]) * inv[:, None]
            t2 = (hi[None] - Oo[:, None]) * inv[:, None]
        tn = np.nanmax(np.minimum(t1, t2), axis=2)
        tf = np.nanmin(np.maximum(t1, t2), axis=2)
        return (tf >= np.maximum(tn, 0.0))

    def sph_hits(Oo, D):
        d = D / np.linalg.norm(D, axis=1, keepdims=True)
        oc = bc[None] - Oo[:, None]
        hh = (oc * d[:, None]).sum(axis=2)
        cc = (oc * oc).sum(axis=2) - br[None] ** 2
        return (hh * hh - cc >= 0) & ((hh >= 0) | (cc <= 0))

    stats = {}
    for _ in range(n_warps):
        x, y = rng.integers(0, w), rng.integers(0, h)
        m = 32
        jit = rng.uniform(-0.5, 0.5, (m, 2))
        sample = p00 + px * (x + jit[:, :1]) + py * (y + jit[:, 1:])
        dd = rng.uniform(-1, 1, (m * 4, 2)); dd = dd[(dd ** 2).sum(1) <= 1][:m]
        Oo = pos + du * dd[:, :1] + dv * dd[:, 1:]
        D = (pos + (sample - pos) * 1.0) - Oo
        alive = np.ones(m, bool)
        for seg in range(8):
            if not alive.any():
                break
            bh = box_hits(Oo[alive], D[alive]); sh = sph_hits(Oo[alive], D[alive])
            s = stats.setdefault(seg, dict(warps=0, lanes=0, box_lane=0, box_union=0, sph_lane=0, sph_union=0, box_max=0))
            s["warps"] += 1; s["lanes"] += int(alive.sum())
            s["box_lane"] += int(bh.sum()); s["box_union"] += int(bh.any(axis=0).sum()); s["box_max"] += int(bh.sum(axis=1).max())
            s["sph_lane"] += int(sh.sum()); s["sph_union"] += int(sh.any(axis=0).sum())
            # closest hit (plain fp64)
            a = (D * D).sum(1)
            oc = c[None] - Oo[:, None]
            hh = (oc * D[:, None]).sum(2)
            cc = (oc * oc).sum(2) - r[None] ** 2
            disc = hh * hh - a[:, None] * cc
            sq = np.sqrt(np.maximum(disc, 0))
            t0 = (hh - sq) / a[:, None]; t1 = (hh + sq) / a[:, None]
            t = np.where(t0 > 1e-6, t0, np.where(t1 > 1e-6, t1, np.inf))
            t = np.where(disc < 0, np.inf, t)
            best = t.argmin(1); bt = t.min(1)
            hit = np.isfinite(bt) & alive
            alive = hit.copy()
            P = Oo + D * bt[:, None]
            N = (P - c[best]) / r[best, None]
            front = (D * N).sum(1) < 0
            N = np.where(front[:, None], N, -N)
            u = rng.normal(size=(m, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
            ud = D / np.linalg.norm(D, axis=1, keepdims=True)
            refl = ud - 2 * (ud * N).sum(1, keepdims=True) * N
            k = kind[best]
            ri = np.where(front, 1 / prm[best, 0], prm[best, 0])
            cos = np.minimum(-(ud * N).sum(1), 1.0)
            perp = (ud + N * cos[:, None]) * ri[:, None]
            par = -np.sqrt(np.abs(1 - (perp ** 2).sum(1)))[:, None] * N
            cannot = ri * np.sqrt(1 - cos * cos) > 1
            D2 = np.where((k == 0)[:, None], N + u, np.where((k == 1)[:, None], refl + u * prm[best, 3:4],
                                                             np.where(cannot[:, None], refl, perp + par)))
            Oo = np.where(hit[:, None], P, Oo); D = np.where(hit[:, None], D2, D)
    print("seg  warps lanes/warp | boxes: per-lane  max-lane  warp-union | bounding spheres: per-lane  warp-union")
    tot = dict(w=0, bu=0, su=0, bl=0, l=0)
    for seg, s in sorted(stats.items()):
        print("%3d %6d %9.1f | %14.2f %9.2f %11.2f | %26.2f %11.2f" % (
            seg, s["warps"], s["lanes"] / s["warps"], s["box_lane"] / s["lanes"], s["box_max"] / s["warps"], s["box_union"] / s["warps"],
            s["sph_lane"] / s["lanes"], s["sph_union"] / s["warps"]))
        tot["w"] += s["warps"]; tot["bu"] += s["box_union"]; tot["su"] += s["sph_union"]; tot["bl"] += s["box_lane"]; tot["l"] += s["lanes"]
    print("all segments: box union per warp-segment %.2f of %d, sphere union %.2f; per lane %.2f" % (
        tot["bu"] / tot["w"], len(clusters), tot["su"] / tot["w"], tot["bl"] / tot["l"]))


if __name__ == "__main__":
    main()

"""Conservativeness of the BVH culling (bvh_closest_hit in tray_kernels.cuh): a node may be skipped only if no sphere inside
it can be accepted by the exact test. Model of the slab test in Python floats (same fp64 operations: 1/d, (plane - o)*inv,
fmin/fmax that drop NaNs, the 1e-9 relative slack, boxes padded by 9.4e-10 relative like both builders do), attacked with
rays that DO hit the sphere (reference Sphere.Hit in oracle/pyref.py): grazing hits near box corners, directions almost
parallel to a slab (1/d overflows to inf), origins on box planes and inside the box, huge and tiny scales."""
import math

import numpy as np


def fmin(a, b):
    return b if a != a else (a if b != b else min(a, b))


def fmax(a, b):
    return b if a != a else (a if b != b else max(a, b))


def padded_box(c, r):
    lo, hi = [], []
    for k in range(3):
        l, h = c[k] - r, c[k] + r
        m = max(abs(l), abs(h)) + 1.0
        lo.append(l - m * 9.4e-10)
        hi.append(h + m * 9.4e-10)
    return lo, hi


def slab_passes(lo, hi, o, d, best_t):
    tn, tf = 0.0, best_t
    for k in range(3):
        inv = math.inf if d[k] == 0 else 1.0 / d[k]
        inv = math.copysign(math.inf, d[k]) if d[k] == 0 else inv
        with np.errstate(invalid="ignore"):
            t1 = float(np.float64(lo[k] - o[k]) * np.float64(inv))
            t2 = float(np.float64(hi[k] - o[k]) * np.float64(inv))
        tn = fmax(tn, fmin(t1, t2))
        tf = fmin(tf, fmax(t1, t2))
    return tn <= tf * 1.000000001 + 1e-300


def test_a_node_holding_a_sphere_that_is_hit_is_never_culled():
    from oracle import pyref
    rs = np.random.RandomState(31)
    hits = 0
    for trial in range(30000):
        scale = float(rs.choice([1e-3, 1.0, 12.0, 1e4]))
        r = float(rs.choice([0.2, 1.0, 0.01])) * scale
        c = tuple(float(v) for v in rs.uniform(-10, 10, 3) * scale)
        lo, hi = padded_box(c, r)
        # the node may be a union with other spheres: grow it at random (culling a bigger box is never less conservative)
        if rs.rand() < 0.5:
            g = rs.uniform(0, 3, 6) * r
            lo = [lo[k] - g[k] for k in range(3)]
            hi = [hi[k] + g[3 + k] for k in range(3)]
        kind = trial % 4
        dirn = rs.normal(0, 1, 3)
        dirn /= np.linalg.norm(dirn)
        if kind == 0:      # grazing hit: passes the centre at distance r*(1 - eps)
            perp = np.cross(dirn, rs.normal(0, 1, 3))
            perp /= np.linalg.norm(perp)
            eps = float(rs.choice([1e-15, 1e-12, 1e-9, 1e-6, 1e-3]))
            o = np.array(c) + perp * r * (1 - eps) - dirn * float(rs.uniform(0.5, 20.0)) * r
        elif kind == 1:    # almost axis-parallel direction, origin on the plane of a face of the box
            ax = int(rs.randint(3))
            dirn = np.zeros(3)
            dirn[(ax + 1) % 3] = 1.0
            dirn[ax] = float(rs.choice([0.0, 1e-300, -1e-300, 1e-18, -1e-18]))
            o = np.array(c) - dirn * 5 * r
            o[ax] = float(rs.choice([lo[ax], hi[ax], c[ax], c[ax] + 0.999999 * r]))
        elif kind == 2:    # origin inside the sphere / the box
            o = np.array(c) + rs.uniform(-0.9, 0.9, 3) * r
        else:
            o = np.array(c) - dirn * float(rs.uniform(1.0, 1e3)) * r + rs.normal(0, 0.5, 3) * r
        o = tuple(float(v) for v in o)
        d = tuple(float(v) for v in dirn * float(rs.choice([1e-3, 1.0, 10.0])))
        if not any(d):
            continue
        hit = pyref.sphere_hit(c, r, o, d, 1e-6, math.inf)
        if hit is None:
            continue
        hits += 1
        t = hit[0]
        for best_t in (math.inf, math.nextafter(t, math.inf), t * (1 + 1e-12), t * 2):
            assert slab_passes(lo, hi, o, d, best_t), (trial, kind, o, d, c, r, t, best_t)
    assert hits > 8000

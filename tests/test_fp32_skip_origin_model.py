"""Model check behind the fp32 fast path's TRAY_FP32_SKIP_ORIGIN (tray_b200/csrc/tray_kernels.cuh): a ray that leaves a sphere to
its outside is not tested against that sphere again.

1. In strict float64 (the oracle's Sphere.Hit, ray/objects.go:81-104) the skipped test never reports a hit: hit points of random
   rays on the ground (r = 1000), the small and the unit spheres of the benchmark scene, scattered outwards like Lambertian.Scatter
   (normal + unit vector, ray/materials.go:13-21) -- no re-hit beyond FrontEpsilon. The rule drops work, not hits.
2. In float32 the same test DOES report a hit now and then (measured on the GPU: 0.8 % of the skipped tests; here: the ground, where
   c = |C-O|^2 - r^2 cancels to about +-0.06 at r = 1000, and mostly for short scattered directions, FrontEpsilon being a ray
   parameter, not a distance). Such a hit is a BACK-face hit (the ray is leaving): the face normal flips inwards, the scattered ray
   enters the sphere and the path bounces inside the ground until the depth limit -- about 35 more segments per event. That is
   where the old fp32 path's 8 % of extra ray segments and its darker speckles came from (2.97 -> 2.74 segments per path, 52 -> 61 dB)."""
import numpy as np
import pytest

from oracle import oracle as O

INF = float("inf")


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for case in range(n):
        kind = case % 3
        if kind == 0:
            c, r = np.array([0.0, -1000.0, 0.0]), 1000.0
            target = np.array([rng.uniform(-11, 11), 0.0, rng.uniform(-11, 11)])
        elif kind == 1:
            c, r = np.array([rng.uniform(-11, 11), 0.2, rng.uniform(-11, 11)]), 0.2
            target = c + rng.normal(0, 0.1, 3)
        else:
            c, r = np.array([rng.uniform(-4, 4), 1.0, rng.uniform(-1, 1)]), 1.0
            target = c + rng.normal(0, 0.5, 3)
        origin = np.array([13.0, 2.0, 3.0]) + rng.normal(0, 0.05, 3)
        ok, t, p, nrm, front = O.sphere_hit(c, r, origin, target - origin, 1e-6, INF)
        if not ok:
            continue
        uv = rng.normal(size=3)
        uv /= np.linalg.norm(uv)
        d2 = nrm + uv
        if np.dot(d2, (p - c) / r) <= 0:
            continue
        yield kind, c, r, p, d2


def test_strict_fp64_never_rehits_the_sphere_it_leaves_outwards():
    n = 0
    for kind, c, r, p, d2 in _cases(9000, 5):
        n += 1
        assert not O.sphere_hit(c, r, p, d2, 1e-6, INF)[0], (kind, c, r, p, d2)
    assert n > 5000


def _sphere_hit_f32(c, r, o, d, tmin):
    """Sphere.Hit in float32, operation by operation (the fp32 kernel's exact test without its fused discriminant): the root or None."""
    f = np.float32
    oc = c - o
    a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    h = d[0] * oc[0] + d[1] * oc[1] + d[2] * oc[2]
    cc = (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2]) - r * r
    disc = h * h - a * cc
    if disc < 0:
        return None
    sq = np.sqrt(disc)
    for root in ((h - sq) / a, (h + sq) / a):
        if root > f(tmin):
            return root
    return None


def test_float32_rehits_the_ground_without_the_rule():
    """The whole bounce in float32: camera ray -> hit point on the ground -> Lambertian direction -> Sphere.Hit of the ground again.
    A reported re-hit is a back-face hit, so the next scattered ray would start inside the ground sphere."""
    f = np.float32
    rng = np.random.default_rng(7)
    c, r = np.array([0.0, -1000.0, 0.0], dtype=f), f(1000.0)
    n = rehits = 0
    for _ in range(6000):
        o = (np.array([13.0, 2.0, 3.0]) + rng.normal(0, 0.05, 3)).astype(f)
        d = (np.array([rng.uniform(-11, 11), 0.0, rng.uniform(-11, 11)]) - o).astype(f)
        t = _sphere_hit_f32(c, r, o, d, 1e-3)
        if t is None:
            continue
        p = o + d * t
        nrm = (p - c) / r
        uv = rng.normal(size=3)
        d2 = nrm + (uv / np.linalg.norm(uv)).astype(f)
        if np.dot(d2, nrm) <= 0:
            continue
        n += 1
        t2 = _sphere_hit_f32(c, r, p, d2, 1e-3)
        if t2 is not None:
            rehits += 1
            p2 = p + d2 * t2
            assert np.dot(d2, (p2 - c) / r) > 0  # FrontFace = false: Lambertian.Scatter would continue INTO the ground
    print("float32 re-hits of the ground beyond 1e-3: %d of %d" % (rehits, n))
    assert n > 4000 and 0.002 * n < rehits < 0.05 * n

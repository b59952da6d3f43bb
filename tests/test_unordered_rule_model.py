"""The BVH visits spheres in tree order, not in Scene.Objects order. bvh_closest_hit therefore uses an order-independent form
of Scene.Hit (exact_test_unordered in tray_kernels.cuh): per sphere the root Sphere.Hit would accept with tmax = +Inf (near
root if it is > tmin, else the far root if that is > tmin), then the minimum by (t, index). This must equal the reference's
scan in slice order with a shrinking interval and "strictly closer wins" -- including exact ties (duplicates), nested shells
and origins inside several spheres. Both forms are evaluated here in Python floats (Go/amd64 semantics)."""
import math

import numpy as np


def unordered_hit(spheres, o, d, tmin):
    from oracle import pyref
    best = None
    order = list(range(len(spheres)))
    np.random.RandomState(len(spheres)).shuffle(order)       # any visiting order
    for i in order:
        center, radius = spheres[i][0], spheres[i][1]
        r = pyref.sphere_hit(center, radius, o, d, tmin, math.inf)
        if r is None:
            continue
        if best is None or r[0] < best[1] or (r[0] == best[1] and i < best[0]):
            best = (i,) + r
    return best


def test_unordered_form_equals_the_reference_scan():
    from oracle import pyref
    rs = np.random.RandomState(21)
    ties = inside = total_hits = 0
    for trial in range(1500):
        n = int(rs.randint(1, 25))
        spheres = []
        for k in range(n):
            c = tuple(float(v) for v in rs.uniform(-2, 2, 3))
            r = float(rs.choice([0.2, 0.5, 1.0, 3.0]))
            spheres.append((c, r, 0, (0.5, 0.5, 0.5, 0.0)))
            if rs.rand() < 0.3:   # exact duplicate (tie -> lowest index) and a concentric shell (back-face hits)
                spheres.append((c, r, 1, (0.9, 0.9, 0.9, 0.0)))
                spheres.append((c, r * 0.9, 2, (1.5, 0.0, 0.0, 0.0)))
        perm = rs.permutation(len(spheres))
        spheres = [spheres[i] for i in perm]
        o = tuple(float(v) for v in rs.uniform(-3, 3, 3))
        if rs.rand() < 0.3:
            o = spheres[0][0]                                  # the centre of a sphere: inside it (and its shell)
            inside += 1
        d = rs.normal(0, 1, 3)
        if rs.rand() < 0.5:
            d = np.subtract(spheres[int(rs.randint(len(spheres)))][0], o) + rs.normal(0, 0.2, 3)
        if not np.any(d):
            continue
        d = tuple(float(v) for v in d)
        want = pyref.scene_hit(spheres, o, d, 1e-6, math.inf)
        got = unordered_hit(spheres, o, d, 1e-6)
        assert (want is None) == (got is None)
        if want is not None:
            total_hits += 1
            assert want[0] == got[0] and want[1] == got[1] and want[2] == got[2] and want[3] == got[3] and want[4] == got[4], trial
            same_t = [i for i, s in enumerate(spheres) if (lambda r: r is not None and r[0] == want[1])(pyref.sphere_hit(s[0], s[1], o, d, 1e-6, math.inf))]
            if len(same_t) > 1:
                ties += 1
                assert want[0] == min(same_t)
    assert total_hits > 800 and ties > 100 and inside > 200

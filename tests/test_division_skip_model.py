"""resolve_candidates (tray_kernels.cuh) skips the true division root = x / a when the quotient certainly falls outside
(tmin, best_t): x >= best_t*a*(1+2^-50)... => fl(x/a) >= best_t, and x <= tmin*a*(1-2^-50)... => fl(x/a) <= tmin. Python
floats are the same IEEE doubles: attack the two claims at the boundaries (x within a few ulps of best_t*a and tmin*a)."""
import math

import numpy as np

UP, DN = 1.0000000000000009, 0.9999999999999991   # the constants of the kernel


def test_skipped_divisions_could_never_have_been_accepted():
    rs = np.random.RandomState(8)
    tmin = 1e-6
    checked_hi = checked_lo = 0
    for _ in range(200000):
        a = float(10.0 ** rs.uniform(-6, 3)) * float(rs.uniform(1, 2))
        best_t = float(10.0 ** rs.uniform(-5, 4)) * float(rs.uniform(1, 2)) if rs.rand() < 0.9 else math.inf
        hi, lo = best_t * a * UP, tmin * a * DN
        # x just around the upper bound
        if math.isfinite(hi):
            x = hi
            for _ in range(int(rs.randint(0, 4))):
                x = math.nextafter(x, math.inf if rs.rand() < 0.5 else -math.inf)
            if x >= hi:                       # the kernel skips: the root must indeed fail `root < best_t`
                checked_hi += 1
                assert not (x / a < best_t), (x, a, best_t)
        # x just around the lower bound
        x = lo
        for _ in range(int(rs.randint(0, 4))):
            x = math.nextafter(x, math.inf if rs.rand() < 0.5 else -math.inf)
        if x <= lo:                           # skipped: the root must indeed fail `root > tmin`
            checked_lo += 1
            assert not (x / a > tmin), (x, a)
    assert checked_hi > 50000 and checked_lo > 50000


def test_the_bounds_are_tight_enough_to_matter():
    """Sanity of the other direction: roots that ARE acceptable are never skipped (x strictly between lo and hi whenever
    tmin < fl(x/a) < best_t)."""
    rs = np.random.RandomState(9)
    tmin = 1e-6
    for _ in range(200000):
        a = float(10.0 ** rs.uniform(-6, 3)) * float(rs.uniform(1, 2))
        best_t = float(10.0 ** rs.uniform(-5, 4)) * float(rs.uniform(1, 2))
        root = float(rs.uniform(tmin, best_t))
        if rs.rand() < 0.5:                   # hug one of the ends
            root = math.nextafter(best_t, 0.0) if rs.rand() < 0.5 else math.nextafter(tmin, 1.0)
        x = root * a
        for _ in range(int(rs.randint(0, 3))):
            x = math.nextafter(x, math.inf if rs.rand() < 0.5 else -math.inf)
        q = x / a
        if tmin < q < best_t:
            assert lo_hi_ok(x, a, best_t, tmin), (x, a, best_t)


def lo_hi_ok(x, a, best_t, tmin):
    return tmin * a * DN < x < best_t * a * UP

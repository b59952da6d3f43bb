"""Exact model of the fp32 pre-filter (filter_scan in tray_b200/csrc/tray_kernels.cuh): every fp32 operation of the device
code is reproduced with correct single rounding (fused multiply-adds through exact rational arithmetic), and the claim the
whole default kernel rests on is attacked where it is weakest -- rays that graze spheres by a relative 1e-9..1e-3, origins on
and inside surfaces, spheres straight behind the origin, large coordinates: whenever the filter says "certainly missed", the
strict fp64 Sphere.Hit of the reference (oracle/pyref.py = Go/amd64 semantics) must return no hit for tmin = 1e-6, tmax = +Inf."""
import math
from fractions import Fraction

import numpy as np

F32 = np.float32
U32 = F32(5.9604645e-8)


def rn32(x):
    """Fraction -> nearest float32 (ties to even), normal range."""
    if x == 0:
        return F32(0.0)
    s = -1 if x < 0 else 1
    x = abs(x)
    e = x.numerator.bit_length() - x.denominator.bit_length()
    if Fraction(2) ** e > x:
        e -= 1
    q = x / Fraction(2) ** (e - 23)            # in [2^23, 2^24)
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and n & 1):
        n += 1
    return F32(s * float(n) * 2.0 ** (e - 23))


def fma32(a, b, c):
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return F32(np.float64(a) * np.float64(b) + np.float64(c))     # inf / nan propagate as in hardware
    return rn32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def sign(x):
    return bool(np.signbit(x)) and not np.isnan(x)                     # a canonical NaN has sign 0 on the device


def ray_constants(o, d, filt_mc, filt_r2max):
    ox, oy, oz = o
    dx, dy, dz = d
    a = dx * dx + dy * dy + dz * dz
    inv_n = 1.0 / math.sqrt(a)
    ddx, ddy, ddz = dx * inv_n, dy * inv_n, dz * inv_n
    fd = (F32(ddx), F32(ddy), F32(ddz))
    ndo = -F32(ddx * ox + ddy * oy + ddz * oz)
    mo = F32(1.0000002) * max(abs(F32(ox)), abs(F32(oy)), abs(F32(oz)))
    R = F32(filt_mc) + mo
    eh = F32(17.5) * U32 * R
    e = F32(1.03) * U32 * (F32(22.0) * R * R + F32(6.2) * F32(filt_r2max)) + F32(1e-30)
    noot = F32(float(e) - (ox * ox + oy * oy + oz * oz))
    if not (mo < F32(1e6)):
        noot, eh = F32(np.inf), F32(0.0)
    ndo = ndo + eh
    p = (F32(2.0 * ox), F32(2.0 * oy), F32(2.0 * oz))
    return fd, p, ndo, noot


def filter_says_missed(consts, c, r):
    fd, p, ndo, noot = consts
    cx, cy, cz = F32(c[0]), F32(c[1]), F32(c[2])
    nk = -F32((c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) - r * r)          # table entry, built on the host in fp64
    h = fma32(fd[0], cx, fma32(fd[1], cy, fma32(fd[2], cz, ndo)))
    nko = nk + noot
    nc = fma32(p[0], cx, fma32(p[1], cy, fma32(p[2], cz, nko)))
    v1 = fma32(h, h, nc)
    return sign(v1) or (sign(h) and sign(nc))


def test_filter_never_skips_a_sphere_the_strict_test_would_hit():
    from oracle import pyref
    rs = np.random.RandomState(11)
    skipped = kept = hits = 0
    for trial in range(20000):
        scale = float(rs.choice([1.0, 1.0, 12.0, 200.0]))
        r = float(rs.choice([0.2, 1.0, 0.05, 3.0])) * min(scale, 12.0) / 1.0 if scale > 1 else float(rs.choice([0.2, 1.0, 0.05]))
        c = tuple(float(v) for v in rs.uniform(-scale, scale, 3))
        filt_mc = float(np.float32(max(abs(v) for v in c)) * np.float32(1.0000002)) + float(rs.choice([0.0, scale]))
        filt_r2max = float(np.float32(r * r) * np.float32(1.0000002)) * float(rs.choice([1.0, 4.0]))
        if filt_mc > 256 or filt_r2max > 256:
            continue                              # such spheres are stored as NaN (always exact-tested)
        kind = trial % 5
        dirn = rs.normal(0, 1, 3)
        dirn /= np.linalg.norm(dirn)
        if kind == 0:      # grazing: the ray passes the centre at distance r * (1 +- eps)
            perp = np.cross(dirn, rs.normal(0, 1, 3))
            perp /= np.linalg.norm(perp)
            eps = float(rs.choice([1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3])) * float(rs.choice([-1, 1]))
            o = np.array(c) + perp * r * (1 + eps) - dirn * float(rs.uniform(0.1, 3.0) * scale)
        elif kind == 1:    # origin on the surface (a scattered ray leaving the sphere it just hit)
            n = rs.normal(0, 1, 3)
            n /= np.linalg.norm(n)
            o = np.array(c) + n * r * (1 + float(rs.choice([0, 1e-15, -1e-15, 1e-9, -1e-9])))
            dirn = n + rs.normal(0, 1, 3) * float(rs.choice([0.01, 1.0]))
        elif kind == 2:    # origin inside
            o = np.array(c) + rs.uniform(-0.5, 0.5, 3) * r
        elif kind == 3:    # sphere straight behind / in front at grazing angle
            o = np.array(c) + dirn * r * float(rs.choice([1.0000001, 1.5, 30.0])) * float(rs.choice([-1, 1]))
            dirn = dirn + rs.normal(0, 1, 3) * float(rs.choice([1e-7, 1e-3, 0.3]))
        else:              # anything
            o = rs.uniform(-1.5, 1.5, 3) * scale
        d = tuple(float(v) for v in dirn * float(rs.choice([1e-3, 1.0, 10.0])))
        o = tuple(float(v) for v in o)
        consts = ray_constants(o, d, filt_mc, filt_r2max)
        hit = pyref.sphere_hit(c, r, o, d, 1e-6, math.inf)
        hits += hit is not None
        if filter_says_missed(consts, c, r):
            skipped += 1
            assert hit is None, (trial, kind, o, d, c, r)
        else:
            kept += 1
    assert skipped > 2500 and hits > 2500 and kept > 2500      # the attack did reach both sides of the filter


def test_fp64_sign_bit_rule_never_drops_a_hit():
    """The all-fp64 loops (and the re-check in front of every exact test) call a sphere "certainly missed" from three sign
    bits: disc < 0, or h < 0 with c >= 0 (miss_bits in tray_device.cuh, strict evaluation order of sphere_terms). Same attack
    as above, in fp64: the reference's Sphere.Hit must return no hit whenever the rule fires -- tangent rays, origins on the
    surface (c = +-0, tiny), h = -0 included."""
    from oracle import pyref
    rs = np.random.RandomState(12)
    fired = hits = 0
    for trial in range(30000):
        scale = float(rs.choice([1.0, 12.0, 1000.0]))
        r = float(rs.choice([0.2, 1.0, 1000.0, 0.01]))
        c = tuple(float(v) for v in rs.uniform(-scale, scale, 3))
        dirn = rs.normal(0, 1, 3)
        dirn /= np.linalg.norm(dirn)
        kind = trial % 4
        if kind == 0:
            perp = np.cross(dirn, rs.normal(0, 1, 3))
            perp /= np.linalg.norm(perp)
            eps = float(rs.choice([0.0, 1e-16, 1e-15, 1e-12, 1e-9, 1e-6])) * float(rs.choice([-1, 1]))
            o = np.array(c) + perp * r * (1 + eps) - dirn * float(rs.uniform(0.0, 3.0) * r)
        elif kind == 1:
            n = rs.normal(0, 1, 3)
            n /= np.linalg.norm(n)
            o = np.array(c) + n * r
            dirn = n * float(rs.choice([1, -1])) + rs.normal(0, 1, 3) * float(rs.choice([0.0, 1e-9, 0.5]))
        elif kind == 2:
            o = np.array(c) + dirn * r * float(rs.choice([-1.0, -1.0000001, -5.0, 1.0, 5.0]))
            if rs.rand() < 0.5:
                dirn = np.array([0.0, 0.0, 1.0]) if rs.rand() < 0.5 else dirn   # axis-parallel: exact zeros in h
        else:
            o = rs.uniform(-1.5, 1.5, 3) * scale
        if not np.any(dirn):
            continue
        o = tuple(float(v) for v in o)
        d = tuple(float(v) for v in dirn * float(rs.choice([1e-3, 1.0, 10.0])))
        ocx, ocy, ocz = c[0] - o[0], c[1] - o[1], c[2] - o[2]
        a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        h = d[0] * ocx + d[1] * ocy + d[2] * ocz
        cc = (ocx * ocx + ocy * ocy + ocz * ocz) - r * r
        disc = h * h - a * cc
        sb = lambda v: math.copysign(1.0, v) < 0                     # the sign BIT, as the kernel reads it
        hit = pyref.sphere_hit(c, r, o, d, 1e-6, math.inf)
        hits += hit is not None
        if sb(disc) or (sb(h) and not sb(cc)):
            fired += 1
            assert hit is None, (trial, o, d, c, r, h, cc, disc)
    assert fired > 5000 and hits > 5000

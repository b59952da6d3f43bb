"""Second, independent restatement of the hot path -- pure Python, written from the reference's Go sources (not from
tray_oracle.c) and only for small cases. TEST INFRASTRUCTURE: it exists so that a systematic mistake in the C oracle
(evaluation order, tie rule, draw order, unwind order) would show up as a bit-level disagreement between two restatements
that share no code. Python floats are IEEE doubles and CPython never fuses a*b+c, i.e. Go/amd64 semantics.

Shared with the C oracle on purpose (third-party arithmetic that is absent from the reference tree): the
ziggurat tables (parsed from zig_tables.h) and the msun log/exp forms of its slow path (called through the C oracle)."""
import math
import os
import re

import numpy as np

_M64 = (1 << 64) - 1
_M128 = (1 << 128) - 1
_HERE = os.path.dirname(os.path.abspath(__file__))


def _tables():
    src = open(os.path.join(_HERE, "zig_tables.h")).read()

    def arr(name):
        body = re.search(name + r"\[128\] = \{(.*?)\};", src, re.S).group(1)
        return [t.strip() for t in body.replace("\n", " ").split(",") if t.strip()]
    kn = [int(t.rstrip("u"), 16) for t in arr("ZIG_KN_INIT")]
    wn = [np.float32(float.fromhex(t.rstrip("f"))) for t in arr("ZIG_WN_INIT")]
    fn = [np.float32(float.fromhex(t.rstrip("f"))) for t in arr("ZIG_FN_INIT")]
    inv_rn = float.fromhex(re.search(r"#define ZIG_INV_RN (\S+)", src).group(1))
    return kn, wn, fn, 3.442619855899, inv_rn


_KN, _WN, _FN, _RN, _INV_RN = _tables()


class Rand:
    """fortio.org/rand over math/rand/v2 PCG-DXSM: rand.NewIdx(idx, seed) (ray/tracer.go:121)."""
    MUL = (2549297995355413924 << 64) | 4865540595714422341
    INC = (6364136223846793005 << 64) | 1442695040888963407

    def __init__(self, idx, seed):
        self.s = ((idx & _M64) << 64) | (seed & _M64)

    def Uint64(self):
        self.s = (self.s * self.MUL + self.INC) & _M128
        hi, lo = self.s >> 64, self.s & _M64
        hi ^= hi >> 32
        hi = (hi * 0xda942042e4dd58b5) & _M64
        hi ^= hi >> 48
        return (hi * (lo | 1)) & _M64

    def Float64(self):
        return float(self.Uint64() & ((1 << 53) - 1)) / 9007199254740992.0   # (u << 11 >> 11) / 2^53

    def NormFloat64(self):
        from . import oracle as O   # msun log / exp of the slow path only
        L = O.lib()
        while True:
            u = self.Uint64()
            j = u & 0xFFFFFFFF
            if j >= 1 << 31:
                j -= 1 << 32
            i = (u >> 32) & 0x7F
            x = float(j) * float(_WN[i])
            if abs(j) < _KN[i]:
                return x
            if i == 0:
                while True:
                    x = -L.oracle_go_log(self.Float64()) * _INV_RN
                    y = -L.oracle_go_log(self.Float64())
                    if y + y >= x * x:
                        break
                return _RN + x if j > 0 else -_RN - x
            lhs = np.float32(_FN[i] + np.float32(self.Float64()) * np.float32(_FN[i - 1] - _FN[i]))
            if lhs < np.float32(L.oracle_go_exp(-.5 * x * x)):
                return x

    def UnitVector(self):
        while True:
            x, y, z = self.NormFloat64(), self.NormFloat64(), self.NormFloat64()
            r = math.sqrt(x * x + y * y + z * z)
            if r > 1e-24:
                return (x / r, y / r, z / r)

    def InDisc(self, radius):
        while True:
            x = 2 * self.Float64() - 1
            y = 2 * self.Float64() - 1
            if x * x + y * y <= 1:
                return radius * x, radius * y


# ---- ray/vec3.go ----------------------------------------------------------------------------------------------------
def add(u, v): return (v[0] + u[0], v[1] + u[1], v[2] + u[2])                       # :25
def sub(u, v): return (u[0] - v[0], u[1] - v[1], u[2] - v[2])                       # :30
def smul(v, t): return (v[0] * t, v[1] * t, v[2] * t)                                # :92
def mul(u, v): return (u[0] * v[0], u[1] * v[1], u[2] * v[2])                        # :97
def sdiv(v, t): return (v[0] / t, v[1] / t, v[2] / t)                                # :102
def dot(u, v): return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]                        # :58
def neg(v): return (-v[0], -v[1], -v[2])
def length(v): return math.sqrt(dot(v, v))
def unit(v): return sdiv(v, length(v))                                               # :117
def near_zero(v): return abs(v[0]) < 1e-8 and abs(v[1]) < 1e-8 and abs(v[2]) < 1e-8  # :128
def reflect(v, n): return sub(v, smul(n, 2 * dot(v, n)))                             # :134


def refract(uv, n, eta):                                                             # :140-145
    cos_theta = min(dot(neg(uv), n), 1.0)
    perp = smul(add(uv, smul(n, cos_theta)), eta)
    par = smul(n, -math.sqrt(abs(1.0 - dot(perp, perp))))
    return add(perp, par)


# ---- ray/objects.go, ray/materials.go ----------------------------------------------------------------------------------
def sphere_hit(center, radius, o, d, tmin, tmax):
    """Sphere.Hit (objects.go:81-104). Returns None or (t, point, normal, front_face)."""
    oc = sub(center, o)
    a = dot(d, d)
    h = dot(d, oc)
    c = dot(oc, oc) - radius * radius
    disc = h * h - a * c
    if disc < 0:
        return None
    sq = math.sqrt(disc)
    root = (h - sq) / a
    if not (tmin < root < tmax):
        root = (h + sq) / a
        if not (tmin < root < tmax):
            return None
    p = add(o, smul(d, root))
    n = sdiv(sub(p, center), radius)
    front = dot(d, n) < 0
    return root, p, (n if front else neg(n)), front


def scene_hit(spheres, o, d, tmin, tmax):
    """Scene.Hit (objects.go:37-46): slice order, the interval shrinks to the closest so far, strictly closer wins."""
    best, closest = None, tmax
    for i, (center, radius, _, _) in enumerate(spheres):
        r = sphere_hit(center, radius, o, d, tmin, closest)
        if r is not None:
            closest = r[0]
            best = (i,) + r
    return best


def reflectance(cosine, ri):                                                         # materials.go:66-71
    r0 = (1 - ri) / (1 + ri)
    r0 = r0 * r0
    x = 1 - cosine
    return r0 + (1 - r0) * (x * ((x * x) * (x * x)))                                 # math.Pow(x, 5): square-and-multiply


def scatter(kind, prm, rng, d_in, p, n, front):
    """Returns None (absorbed) or (attenuation, new direction); draw order as in materials.go:13-64."""
    if kind == 0:
        d = add(n, rng.UnitVector())
        if near_zero(d):
            d = n
        return (prm[0], prm[1], prm[2]), d
    if kind == 1:
        refl = reflect(unit(d_in), n)
        if prm[3] > 0:
            refl = add(refl, smul(rng.UnitVector(), prm[3]))
        return ((prm[0], prm[1], prm[2]), refl) if dot(refl, n) > 0 else None
    ri = prm[0]
    ratio = 1.0 / ri if front else ri
    u = unit(d_in)
    cos_theta = min(dot(neg(u), n), 1.0)
    sin_theta = math.sqrt(1.0 - cos_theta * cos_theta)
    if ratio * sin_theta > 1.0 or reflectance(cos_theta, ratio) > rng.Float64():     # short-circuit: no draw when it cannot refract
        d = reflect(u, n)
    else:
        d = refract(u, n, ratio)
    return (1.0, 1.0, 1.0), d


def ray_color(spheres, bg_a, bg_b, rng, o, d, depth):
    """Scene.RayColor (objects.go:49-62), recursive like the reference: Mul(attenuation, RayColor(scattered, depth-1))."""
    if depth <= 0:
        return (0.0, 0.0, 0.0)
    hit = scene_hit(spheres, o, d, 1e-6, math.inf)
    if hit is None:
        u = unit(d)
        a = 0.5 * (u[1] + 1.0)
        return add(smul(bg_a, 1.0 - a), smul(bg_b, a))                               # AmbientLight.Hit, objects.go:64-73
    i, t, p, n, front = hit
    s = scatter(spheres[i][2], spheres[i][3], rng, d, p, n, front)
    if s is None:
        return (0.0, 0.0, 0.0)
    return mul(s[0], ray_color(spheres, bg_a, bg_b, rng, p, s[1], depth - 1))


# ---- ray/camera.go:113-142, ray/tracer.go:85-155 ----------------------------------------------------------------------------
def get_ray(cam, rng, px, py, ox, oy):
    """Camera.GetRay on an initialised camera (fields of oracle.Camera / tray_camera)."""
    pos = tuple(cam.position)
    sample = add(add(tuple(cam.pixel00), smul(tuple(cam.pixel_x), px + ox)), smul(tuple(cam.pixel_y), py + oy))
    d = sub(sample, pos)
    if cam.aperture > 0:
        dx, dy = rng.InDisc(1.0)
        offset = add(smul(tuple(cam.defocus_u), dx), smul(tuple(cam.defocus_v), dy))
        focus_time = cam.focus_distance / cam.focal_length
        focus_point = add(pos, smul(d, focus_time))
        o = add(pos, offset)
        return o, sub(focus_point, o)
    return pos, d


def render_lines(spheres, bg_a, bg_b, cam, width, spp, max_depth, ray_radius, seed, idx, y0, y1, to_srgb, per_sample=False):
    """Tracer.RenderLines (tracer.go:120-155): ONE stream rand.NewIdx(idx, seed) for the call (or, with per_sample, the
    convention of the throughput mode: one stream per sample, idx = (y*W+x)*spp+s). Returns {(x, y): (r, g, b)} in 8 bits."""
    rng = Rand(idx, seed)
    div = 1.0 / float(spp)
    out = {}
    for y in range(y0, y1):
        for x in range(width):
            csum = (0.0, 0.0, 0.0)
            for s in range(spp):
                if per_sample:
                    rng = Rand((y * width + x) * spp + s, seed)
                ox = oy = 0.0
                if spp > 1:
                    ox, oy = rng.InDisc(ray_radius)
                o, d = get_ray(cam, rng, float(x), float(y), ox, oy)
                csum = add(csum, ray_color(spheres, bg_a, bg_b, rng, o, d, max_depth))
            c = smul(csum, div)
            out[(x, y)] = tuple(to_srgb(v) for v in c)
    return out


def render(spheres, bg_a, bg_b, cam, width, height, spp, max_depth, ray_radius, seed, num_workers, to_srgb):
    """Tracer.Render's fan-out (tracer.go:85-116): one stream idx 0 for a single worker, else chunks of max(4, h/(4W)) rows
    whose stream index is their first row."""
    if num_workers == 1:
        return render_lines(spheres, bg_a, bg_b, cam, width, spp, max_depth, ray_radius, seed, 0, 0, height, to_srgb)
    chunk = max(4, height // (num_workers * 4))
    out = {}
    for y in range(0, height, chunk):
        out.update(render_lines(spheres, bg_a, bg_b, cam, width, spp, max_depth, ray_radius, seed, y, y, min(y + chunk, height), to_srgb))
    return out

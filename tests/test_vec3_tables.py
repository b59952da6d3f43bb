"""The reference's operator tables (ray/vec3_test.go, ray/ray_test.go) replayed against the oracle's restatement of the
Vec3 helpers (what every kernel's arithmetic is checked against) and against the Python host mirror (tray_b200/ray.py,
used by Camera.Initialize). Exact where the reference is exact, 1e-9 where it is (Reflect), properties where it checks
properties (Refract)."""
import math

import pytest

from tray_b200 import ray


def both(name, *args, **kw):
    """(oracle result, host-mirror result or None)"""
    from oracle import oracle as O
    o = O.vec_op(name, *args, **kw)
    host = getattr(ray, name, None)
    h = None
    if host is not None and name not in ("Refract", "Reflect", "Minus", "Surrounds"):
        h = host(*[a for a in args if a is not None], *([kw["t"]] if "t" in kw else []))
    return o, h


def test_add_sub_neg(O):                                   # vec3_test.go:11-34
    o, h = both("Add", (1, 2, 3), (4, 5, 6))
    assert o == (5, 7, 9) == h
    o, h = both("Sub", (5, 7, 9), (1, 2, 3))
    assert o == (4, 5, 6) == h
    n, hn = both("Neg", (1, 2, 3))
    assert n == (-1, -2, -3) == hn and O.vec_op("Add", (5, 7, 9), n) == (4, 5, 6)


@pytest.mark.parametrize("u,vs,want", [((1, 2, 3), [(4, 5, 6)], (5, 7, 9)), ((1, 2, 3), [(4, 5, 6), (7, 8, 9)], (12, 15, 18)),
                                        ((1, 1, 1), [(1, 0, 0), (0, 1, 0), (0, 0, 1)], (2, 2, 2)), ((5, 10, 15), [], (5, 10, 15)),
                                        ((10, 10, 10), [(-5, 0, 5), (3, -3, 0)], (8, 7, 15))])
def test_add_multiple(O, u, vs, want):                      # vec3_test.go:36-58 (left-to-right accumulation)
    acc = u
    for v in vs:
        acc = O.vec_op("Add", acc, v)
    assert tuple(acc) == want


def test_minus_is_u_minus_the_sum(O):                       # vec3.go:44-55, vec3_test.go:60-106: u - (v0 + v1), not (u - v0) - v1
    u, v, w = (10.0, 0.1, 1e16), (3.0, 0.2, 1.0), (2.0, 0.3, 1.0)
    got = O.vec_op("Minus", u, v, w)
    assert got == tuple(a - (b + c) for a, b, c in zip(u, v, w))
    assert got[2] != (u[2] - v[2]) - w[2]                   # the two orders differ in the last bit: the camera set-up depends on it


def test_smul_mul_sdiv_dot(O):                              # vec3_test.go:160-199,317-325
    assert both("SMul", (1, 2, 3), t=2) == ((2, 4, 6), (2, 4, 6))
    assert both("Mul", (1, 2, 3), (4, 5, 6)) == ((4, 10, 18), (4, 10, 18))
    assert both("SDiv", (2, 4, 6), t=2) == ((1, 2, 3), (1, 2, 3))
    assert both("Dot", (1, 2, 3), (4, 5, 6)) == (32.0, 32)
    assert O.vec_op("SDiv", (1, 1, 1), t=3)[0] == 1 / 3     # true division, not multiplication by a reciprocal


@pytest.mark.parametrize("v,want", [((3, 4, 0), 5.0), ((1, 0, 0), 1.0), ((0, 0, 0), 0.0), ((1, 1, 1), math.sqrt(3)), ((-3, -4, 0), 5.0)])
def test_length(O, v, want):                                # vec3_test.go:200-220
    o, h = both("Length", v)
    assert o == want == h
    assert O.vec_op("LengthSquared", v) == v[0] * v[0] + v[1] * v[1] + v[2] * v[2]


@pytest.mark.parametrize("v", [(3, 4, 0), (1, 0, 0), (0, 5, 0), (1, 1, 1), (13, 2, 3)])
def test_unit_is_three_true_divisions(O, v):                # vec3_test.go:222-240, vec3.go:117-120
    o, h = both("Unit", v)
    l = math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    assert o == (v[0] / l, v[1] / l, v[2] / l) == h
    assert abs(O.vec_op("Length", o) - 1) < 1e-15


def test_cross(O):                                          # vec3.go:71-77 (camera basis)
    assert both("Cross", (1, 0, 0), (0, 1, 0)) == ((0, 0, 1), (0, 0, 1))
    assert both("Cross", (0, 1, 0), (13, 2, 3)) == ((3, 0, -13), (3, 0, -13))


@pytest.mark.parametrize("v,want", [((0, 0, 0), True), ((1e-9, 1e-10, 1e-11), True), ((1e-7, 0, 0), False), ((1, 2, 3), False),
                                     ((1e-9, 1e-9, 1e-6), False), ((-1e-10, -1e-11, -1e-12), True), ((1e-10, -1e-11, 1e-12), True)])
def test_near_zero(O, v, want):                             # vec3_test.go:764-787
    o, h = both("NearZero", v)
    assert bool(o) == want == h


@pytest.mark.parametrize("v,n,want", [((1, -1, 0), (0, 1, 0), (1, 1, 0)), ((1, 1, 0), (1, 0, 0), (-1, 1, 0)), ((0, -1, 0), (0, 1, 0), (0, 1, 0)),
                                       ((1 / math.sqrt(2), -1 / math.sqrt(2), 0), (0, 1, 0), (1 / math.sqrt(2), 1 / math.sqrt(2), 0))])
def test_reflect(O, v, n, want):                            # vec3_test.go:789-835 (1e-9)
    got = O.vec_op("Reflect", v, n)
    assert all(abs(g - w) <= 1e-9 for g, w in zip(got, want))


@pytest.mark.parametrize("uv,n,eta,check,decreases", [((0, -1, 0), (0, 1, 0), 1.5, False, False),
                                                        ((1 / math.sqrt(2), -1 / math.sqrt(2), 0), (0, 1, 0), 1.0 / 1.5, True, True),
                                                        ((1 / math.sqrt(2), -1 / math.sqrt(2), 0), (0, 1, 0), 1.5, True, False)])
def test_refract(O, uv, n, eta, check, decreases):          # vec3_test.go:837-904
    r = O.vec_op("Refract", uv, n, t=eta)
    assert not bool(O.vec_op("NearZero", r))
    if check:
        ru = O.vec_op("Unit", r)
        inc = math.acos(abs(O.vec_op("Dot", uv, n)))
        ref = math.acos(min(1.0, abs(O.vec_op("Dot", ru, n))))
        assert (ref < inc) if decreases else (ref > inc)


@pytest.mark.parametrize("lo,hi,t,want", [(0, 10, 5, True), (0, 10, 0, False), (0, 10, 10, False), (0, 10, -1, False), (0, 10, 11, False),
                                           (0, 1, 0.5, True), (0, 1, 0, False), (0, 1, 1, False), (math.inf, -math.inf, 0, False),
                                           (-math.inf, math.inf, 999999, True), (-math.inf, math.inf, -math.inf, False),
                                           (-math.inf, math.inf, math.inf, False), (1e-6, math.inf, 1e-6, False), (1e-6, math.inf, 2e-6, True)])
def test_interval_surrounds_is_exclusive(O, lo, hi, t, want):   # vec3_test.go:388-419; FrontEpsilon = (1e-6, +Inf), vec3.go:218
    assert bool(O.vec_op("Surrounds", (lo, hi, 0), t=t)) == want


def test_sphere_hit_honours_the_exclusive_interval(O):      # the only Interval on the hot path (objects.go:92-97)
    # unit sphere at the origin seen from z = 3: roots t = 2 and t = 4
    hit, t, p, n, front = O.sphere_hit((0, 0, 0), 1.0, (0, 0, 3), (0, 0, -1), 0.0, math.inf)
    assert hit and t == 2.0 and front and tuple(p) == (0, 0, 1) and tuple(n) == (0, 0, 1)
    assert not O.sphere_hit((0, 0, 0), 1.0, (0, 0, 3), (0, 0, -1), 2.0, 4.0)[0]       # both roots sit ON the bounds: rejected
    hit, t, p, n, front = O.sphere_hit((0, 0, 0), 1.0, (0, 0, 3), (0, 0, -1), 2.0, 4.0000001)
    assert hit and t == 4.0 and not front                                             # near root excluded, far root inside: back face


@pytest.mark.parametrize("o,d,t,want", [((0, 0, 0), (1, 0, 0), 0, (0, 0, 0)), ((0, 0, 0), (1, 0, 0), 1, (1, 0, 0)), ((1, 2, 3), (1, 1, 1), 2, (3, 4, 5)),
                                         ((1, 2, 3), (0, 1, 0), -1, (1, 1, 3)), ((2, 3, 4), (0.5, 0.25, 2), 2, (3, 3.5, 8))])
def test_ray_at(O, o, d, t, want):                          # ray_test.go:20-58: At(t) = Origin + Direction*t
    assert O.vec_op("Add", o, O.vec_op("SMul", d, t=t)) == want


# ---- Camera.GetRay (ray/camera_test.go:105-243) against the oracle's restatement ---------------------------------------
def test_get_ray_pinhole_origin_is_the_position(O):         # camera_test.go:105-130
    cam = O.camera_init(100, 100, position=(0, 0, 5), look_at=(0, 0, 0), aperture=0.0, focal_length=1.0)
    r = O.get_rays(cam, [(50, 50, 0, 0), (50, 50, 0, 0), (25, 75, 0, 0)])
    assert (r[:, :3] == [0, 0, 5]).all() and (r[0] == r[1]).all() and (r[0, 3:] != r[2, 3:]).any()


def test_get_ray_depth_of_field_samples_the_lens(O):        # camera_test.go:132-162
    cam = O.camera_init(100, 100, position=(0, 0, 5), look_at=(0, 0, 0), aperture=0.5, focal_length=1.0, focus_distance=5.0)
    r = O.get_rays(cam, [(50, 50, 0, 0)] * 64)
    d = ((r[:, :3] - [0, 0, 5]) ** 2).sum(axis=1) ** 0.5
    assert (r[0, :3] != r[1, :3]).any() and (d <= 0.25).all() and d.max() > 0.05
    # all rays of a pixel meet on the focus plane: origin + dir hits the same point (GetRay aims at pos + dir*focusTime)
    focus = r[:, :3] + r[:, 3:]
    assert abs(focus - focus[0]).max() < 1e-12


def test_get_ray_offsets_move_the_direction_not_the_origin(O):   # camera_test.go:218-243
    cam = O.camera_init(10, 10, position=(0, 0, 0), look_at=(0, 0, -1), vfov=90.0, focal_length=1.0)
    r = O.get_rays(cam, [(5, 5, 0, 0), (5, 5, 0.3, 0.2)])
    assert (r[0, :3] == r[1, :3]).all() and (r[0, 3:] != r[1, 3:]).any()


def test_focus_distance_defaults_to_focal_length():         # camera_test.go:164-175 (host mirror: Initialize stays on the host)
    c = ray.Camera(FocalLength=2.5)
    c.Initialize(100, 100)
    assert c.FocusDistance == c.FocalLength == 2.5

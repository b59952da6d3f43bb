"""Two restatements that share no code -- the C oracle and oracle/pyref.py (pure Python, written from the Go sources) --
must agree bit for bit on the generators, Sphere.Hit, the three Scatter bodies and whole RayColor paths."""
import math

import numpy as np
import pytest


def _spheres(flat):
    return [((flat.cx[i], flat.cy[i], flat.cz[i]), float(flat.r[i]), int(flat.kind[i]), tuple(float(x) for x in flat.params[i])) for i in range(flat.n)]


def test_generators_agree(O):
    from oracle import pyref
    for idx, seed in ((0, 2), (5, 42), (2 ** 40 + 17, 7)):
        r = pyref.Rand(idx, seed)
        assert [r.Uint64() for _ in range(64)] == [int(v) for v in O.rng_u64(idx, seed, 64)]
        r = pyref.Rand(idx, seed)
        assert [r.Float64() for _ in range(64)] == list(O.rng_f64(idx, seed, 64))
        r = pyref.Rand(idx, seed)
        assert [r.NormFloat64() for _ in range(3000)] == list(O.rng_norm(idx, seed, 3000))     # ~35 slow-path draws among them
        r = pyref.Rand(idx, seed)
        assert [r.UnitVector() for _ in range(200)] == [tuple(v) for v in O.rng_unit_vectors(idx, seed, 200)]
        r = pyref.Rand(idx, seed)
        assert [r.InDisc(0.5) for _ in range(200)] == [tuple(v) for v in O.rng_in_disc(idx, seed, 0.5, 200)]


def test_sphere_hit_agrees_on_random_cases(O):
    from oracle import pyref
    rs = np.random.RandomState(1)
    hits = 0
    for k in range(3000):
        c = tuple(rs.uniform(-3, 3, 3))
        rad = float(rs.choice([0.2, 1.0, 1000.0, 0.01]))
        o = tuple(rs.uniform(-4, 4, 3))
        d = tuple(rs.normal(0, 1, 3) * rs.choice([1e-3, 1.0, 10.0]))
        if k % 3 == 0:  # aim at the sphere so that hits, grazing hits and inside origins all occur
            d = tuple(np.subtract(c, o) + rs.normal(0, rad * 0.7, 3))
        tmin, tmax = 1e-6, float(rs.choice([math.inf, 5.0, 0.5]))
        got = pyref.sphere_hit(c, rad, o, d, tmin, tmax)
        ok, t, p, n, front = O.sphere_hit(c, rad, o, d, tmin, tmax)
        assert ok == (got is not None)
        if ok:
            hits += 1
            assert got[0] == t and got[1] == tuple(p) and got[2] == tuple(n) and got[3] == front
    assert hits > 500


@pytest.mark.parametrize("which", ["rich", "default"])
def test_ray_color_paths_agree_bit_for_bit(O, which):
    from oracle import pyref
    from tray_b200 import ray
    if which == "rich":
        flat = O.rich_scene(2)
        cam = O.camera_init(64, 36, **O.RICH_CAMERA)
        bg = ((1.0, 1.0, 1.0), (0.4, 0.65, 1.0))
    else:
        f = ray.DefaultScene().flatten()
        flat = O.FlatScene(f["cx"], f["cy"], f["cz"], f["r"], f["kind"], f["params"], f["bg_a"], f["bg_b"])
        cam = O.camera_init(64, 36, position=(-2, 2, 1), look_at=(0, 0, -1), vfov=20.0)
        bg = (tuple(f["bg_a"]), tuple(f["bg_b"]))
    spheres = _spheres(flat)
    rays = O.get_rays(cam, [(x, y, 0.1, -0.2) for y in range(0, 36, 5) for x in range(0, 64, 7)], idx=3, seed=9)
    nonsky = 0
    for k, r in enumerate(rays):
        o, d = tuple(r[:3]), tuple(r[3:])
        want = O.ray_color(flat, o, d, 50, idx=100 + k, seed=7)
        got = pyref.ray_color(spheres, bg[0], bg[1], pyref.Rand(100 + k, 7), o, d, 50)
        assert got == tuple(want), k
        nonsky += pyref.scene_hit(spheres, o, d, 1e-6, math.inf) is not None
    assert nonsky > 20


@pytest.mark.parametrize("workers,spp,aperture", [(1, 3, 0.1), (2, 2, 0.0), (3, 1, 0.1)])
def test_render_fan_out_and_sample_loop_agree(O, workers, spp, aperture):
    """Tracer.Render / RenderLines: stream per chunk, InDisc(RayRadius) iff spp > 1, aperture draw iff Aperture > 0, sum in
    sample order, * (1/N), ToSRGBA -- the image of the pure-Python restatement equals the C oracle's, pixel for pixel."""
    from oracle import pyref
    w, h, depth, seed = 14, 11, 12, 5
    flat = O.rich_scene(2)
    rc = dict(O.RICH_CAMERA)
    rc["aperture"] = aperture
    cam = O.camera_init(w, h, **rc)
    spheres = _spheres(flat)
    bg = ((1.0, 1.0, 1.0), (0.4, 0.65, 1.0))
    p = O.make_params(w, h, spp=spp, max_depth=depth, seed=seed, num_workers=workers, stream_mode=0)
    want, _, _ = O.render(flat, cam, p)
    got = pyref.render(spheres, bg[0], bg[1], cam, w, h, spp, depth, 0.5, seed, workers, O.linear_to_srgb)
    for (x, y), rgb in got.items():
        assert tuple(int(v) for v in want[y, x, :3]) == rgb, (x, y)
    # the per-sample stream convention of the throughput mode
    p1 = O.make_params(w, h, spp=spp, max_depth=depth, seed=seed, num_workers=1, stream_mode=1)
    want1, _, _ = O.render(flat, cam, p1)
    got1 = pyref.render_lines(spheres, bg[0], bg[1], cam, w, spp, depth, 0.5, seed, 0, 0, h, O.linear_to_srgb, per_sample=True)
    for (x, y), rgb in got1.items():
        assert tuple(int(v) for v in want1[y, x, :3]) == rgb, (x, y)

"""Pins the CPU oracle against every golden vector / known-answer the reference holds for the hot path
(SURVEY.md 8c, Appendix A), and re-expresses the reference's unit-test tables."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN


# ---- third-party RNG (Go math/rand/v2 via fortio.org/rand) ------------------------------------
def test_pcg_known_answer(O):
    # Go's own PCG test vector: NewPCG(1, 2) -> first three Uint64 outputs
    got = [int(v) for v in O.rng_u64(1, 2, 3)]
    assert got == [0xc4f5a58656eef510, 0x9dcec3ad077dec6c, 0xc8d04605312f8088]


def test_float64_range_and_resolution(O):
    f = O.rng_f64(0, 42, 100000)
    assert f.min() >= 0.0 and f.max() < 1.0
    assert np.all(f * 2.0 ** 53 == np.floor(f * 2.0 ** 53))  # 53-bit grid
    assert abs(f.mean() - 0.5) < 0.005


def test_seed7_gives_486_objects(O):
    # benchmark/benchmark.go:42 "We get 486 objects like the c++ version with seed 7"; under
    # NewPCG(uint64(idx)=0, seed) 7 is the FIRST seed >= 1 to give 486.
    counts = [O.rich_scene(s).n for s in range(1, 13)]
    assert counts == [485, 485, 485, 485, 485, 485, 486, 484, 484, 486, 485, 486]


def test_seed2_material_mix(O):
    sc = O.rich_scene(2)
    assert sc.n == 485
    assert np.bincount(sc.kind).tolist() == [401, 56, 28]  # ground+399+big, 55+big, 27+big
    assert (sc.cx[0], sc.cy[0], sc.cz[0], sc.r[0]) == (0.0, -1000.0, 0.0, 1000.0)
    assert sc.r[-3:].tolist() == [1.0, 1.0, 1.0] and sc.kind[-3:].tolist() == [2, 0, 1]


def test_scene_matches_example_png(O):
    """example.png (README.md:30-31: -r 64 -s 8 -d 50 -seed 2 at 1280x720) is the reference's only golden
    artefact of the whole path. Its RNG streams cannot be reproduced pixel-for-pixel, but
    the scene layout, camera, sky and sRGB store can: 16x16 block means must agree to noise level, and
    must NOT agree for a different seed."""
    z = np.load(os.path.join(GOLDEN, "example_png_blocks.npz"))
    blocks = z["blocks"].astype(np.float64) / 16.0
    w, h = int(z["width"]), int(z["height"])
    cam = O.camera_init(w, h, **O.RICH_CAMERA)

    def block_rms(seed):
        p = O.make_params(w, h, spp=2, max_depth=50, seed=2, num_workers=8, stream_mode=1)
        img, _, _ = O.render(O.rich_scene(seed), cam, p)
        b = img[:, :, :3].astype(np.float64).reshape(h // 16, 16, w // 16, 16, 3).mean(axis=(1, 3))
        return float(np.sqrt(((b - blocks) ** 2).mean()))

    good, bad = block_rms(2), block_rms(3)
    assert good < 6.0, good   # 2 rays/pixel noise: ~4.0 (2.0 at 4 rays/pixel); a wrong scene gives ~44
    assert bad > 5 * good, (good, bad)
    # sky rows are RNG-insensitive: exact 8-bit match of the top block row
    p = O.make_params(w, h, spp=1, max_depth=50, seed=2, num_workers=8, stream_mode=1)
    img, _, _ = O.render(O.rich_scene(2), cam, p)
    top = img[:16, :, :3].astype(np.float64).reshape(1, 16, w // 16, 16, 3).mean(axis=(1, 3))
    assert np.abs(top - blocks[:1]).max() < 0.51


def test_ziggurat_table_heads():
    # literals of Go's math/rand/v2 normal.go, checked inside the generator too
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "oracle", "zig_tables.h")).read()
    assert "0x76ad2212u, 0x00000000u, 0x600f1b53u, 0x6ce447a6u, 0x725b46a2u" in src
    assert "0x77d664e5u" in src
    assert open(os.path.join(root, "tray_b200", "csrc", "zig_tables.h")).read() == src


def test_norm_float64_distribution(O):
    n = O.rng_norm(3, 42, 400000)
    assert abs(n.mean()) < 0.01 and abs(n.var() - 1.0) < 0.01
    assert abs((np.abs(n) > 3.442619855899).mean() - 5.76e-4) < 1.5e-4  # tail strip is exercised
    assert abs(np.mean(n ** 4) - 3.0) < 0.06


def test_unit_vector_statistics(O):
    # ray/vec3_test.go:505-646: |len-1| < 1e-9; mean < 0.015; var ~ 1/3 +- 0.01; octants within 15 %
    v = O.rng_unit_vectors(0, 42, 100000)
    assert np.abs(np.linalg.norm(v, axis=1) - 1).max() < 1e-9
    assert np.abs(v.mean(axis=0)).max() < 0.015
    assert np.abs(v.var(axis=0) - 1 / 3).max() < 0.01
    octant = (v[:, 0] > 0) * 4 + (v[:, 1] > 0) * 2 + (v[:, 2] > 0)
    cnt = np.bincount(octant, minlength=8)
    assert np.abs(cnt / (len(v) / 8) - 1).max() < 0.15


def test_in_disc(O):
    d = O.rng_in_disc(1, 42, 0.5, 50000)
    r = np.hypot(d[:, 0], d[:, 1])
    assert r.max() <= 0.5 and abs((r < 0.25).mean() - 0.25) < 0.01  # uniform over the disc


def test_go_math_restatements(O):
    L = O.lib()
    for fov in (20.0, 30.0, 40.0, 60.0, 90.0):
        th = fov * (math.pi / 180.0) / 2.0
        assert L.oracle_go_tan(th) == pytest.approx(math.tan(th), rel=3e-16)
    assert L.oracle_go_tan(20.0 * (math.pi / 180.0) / 2.0) == math.tan(20.0 * (math.pi / 180.0) / 2.0)
    xs = np.random.default_rng(0).random(2000)
    for x in xs:
        assert L.oracle_go_log(x) == pytest.approx(math.log(x), rel=4e-16, abs=1e-300)
        assert L.oracle_go_exp(-6 * x) == pytest.approx(math.exp(-6 * x), rel=4e-16)


# ---- tcolor.LinearToSrgb, ray/vec3_test.go:264-289 -------------------------------------------------
@pytest.mark.parametrize("x,want", [(0.0, 0), (1.0, 255), (0.5, 188), (-0.5, 0), (1.5, 255), (0.25, 137), (0.75, 225)])
def test_linear_to_srgb_table(O, x, want):
    assert O.linear_to_srgb(x) == want


def test_linear_to_srgb_monotone(O):
    xs = np.linspace(0, 1, 20001)
    v = np.array([O.linear_to_srgb(x) for x in xs])
    assert np.all(np.diff(v.astype(int)) >= 0) and v[0] == 0 and v[-1] == 255
    # sky pixels of example.png probed in SURVEY App. A.5: white*(1-a)+blue*a through sRGB, e.g. (217,234,255)
    assert [O.linear_to_srgb(c) for c in (0.6939, 0.8215, 1.0)] == [217, 234, 255]


# ---- Sphere.Hit tables, ray/objects_test.go:48-207 -------------------------------------------------
@pytest.mark.parametrize("fma", [0, 1])
def test_sphere_hit_tables(O, fma):
    INF = float("inf")
    ok, t, p, n, front = O.sphere_hit((0, 0, -5), 1.0, (0, 0, 0), (0, 0, -1), 1e-6, INF, fma)
    assert ok and abs(t - 4.0) < 1e-10 and abs(np.linalg.norm(p - np.array([0, 0, -5.0])) - 1.0) < 1e-10 and front
    ok, *_ = O.sphere_hit((0, 0, -5), 1.0, (0, 0, 0), (0, 1, 0), 1e-6, INF, fma)
    assert not ok
    ok, t, p, n, front = O.sphere_hit((0, 0, 0), 1.0, (5, 0, 0), (-1, 0, 0), 1e-6, INF, fma)
    assert ok and np.linalg.norm(n - np.array([1.0, 0, 0])) < 1e-10 and front
    ok, t, p, n, front = O.sphere_hit((0, 0, 0), 1.0, (0, 0, 0), (0, 0, -1), 0.0, INF, fma)  # from inside
    assert ok and not front and abs(t - 1.0) < 1e-10 and np.allclose(n, (0, 0, 1))
    for lo, hi in ((0.0, 3.0), (10.0, 20.0)):  # hit at t=4 (and 6) excluded by the interval
        ok, *_ = O.sphere_hit((0, 0, -5), 1.0, (0, 0, 0), (0, 0, -1), lo, hi, fma)
        assert not ok
    # Surrounds is strict: a root exactly at tmax is rejected, so ties keep the earlier object
    ok, t, *_ = O.sphere_hit((0, 0, -5), 1.0, (0, 0, 0), (0, 0, -1), 1e-6, 4.0, fma)
    assert not ok


def test_scene_hit_closest_and_ties(O):
    # two spheres, closest t ~ 0.5 wins regardless of order (objects_test.go:182-207)
    def scene(centers):
        c = np.array(centers, dtype=float)
        return O.FlatScene(c[:, 0], c[:, 1], c[:, 2], np.full(len(c), 0.5), np.zeros(len(c), np.uint8), np.zeros((len(c), 4)))
    cam = O.camera_init(1, 1, position=(0, 0, 0), look_at=(0, 0, -1))
    for centers in ([(0, 0, -1), (0, 0, -3)], [(0, 0, -3), (0, 0, -1)]):
        ids, t, n, f = O.first_hit(scene(centers), cam, 1, 1)
        assert abs(t[0, 0] * np.linalg.norm([0, 0, -1]) - 0.5) < 1e-9
    # identical spheres: the tie resolves to the lowest index
    ids, *_ = O.first_hit(scene([(0, 0, -2), (0, 0, -2), (0, 0, -2)]), cam, 1, 1)
    assert ids[0, 0] == 0


# ---- Material.Scatter, ray/materials_test.go -------------------------------------------------------
def test_lambertian_scatter(O):
    for idx in range(20):
        did, att, o, d, draws = O.scatter(0, (0.8, 0.3, 0.3, 0), idx, 42, (0, 0, 0), (0, 0, -1), (1, 2, 3), (0, 0, 1), True)
        assert did and att.tolist() == [0.8, 0.3, 0.3] and o.tolist() == [1, 2, 3] and draws >= 3
        assert np.linalg.norm(d - np.array([0, 0, 1.0])) <= 1 + 1e-9  # normal + unit vector


def test_metal_scatter_and_fuzz_absorb(O):
    did, att, o, d, draws = O.scatter(1, (0.8, 0.8, 0.8, 0.0), 0, 42, (0, 1, 0), (1, -1, 0), (0, 0, 0), (0, 1, 0), True)
    assert did and draws == 0 and att.tolist() == [0.8, 0.8, 0.8]
    s = 1 / math.sqrt(2)
    assert np.allclose(d, (s, s, 0), atol=1e-15)  # mirror reflection of the unit direction
    res = [O.scatter(1, (0.8, 0.8, 0.8, 1.5), i, 42, (0, 1, 0), (1, -0.1, 0), (0, 0, 0), (0, 1, 0), True) for i in range(200)]
    assert any(not r[0] for r in res) and any(r[0] for r in res)  # fuzz 1.5 can absorb (materials_test.go:82-111)
    assert all(r[4] >= 3 for r in res)


def test_dielectric_scatter(O):
    for idx in range(50):
        for front, prm in ((True, 1.5), (False, 1.5), (True, 1 / 1.5)):
            did, att, o, d, draws = O.scatter(2, (prm, 0, 0, 0), idx, 42, (0, 1, 0), (0.3, -1, 0.1), (0, 0, 0), (0, 1, 0), front)
            assert did and att.tolist() == [1.0, 1.0, 1.0] and draws in (0, 1)
    # total internal reflection draws nothing (short-circuit at materials.go:57)
    did, att, o, d, draws = O.scatter(2, (1.5, 0, 0, 0), 0, 42, (0, 1, 0), (1, -0.05, 0), (0, 0, 0), (0, 1, 0), False)
    assert did and draws == 0 and d[1] > 0


def test_reflectance_is_schlick(O):
    for cos in np.linspace(0, 1, 11):
        for ri in (1.5, 1 / 1.5, 1.0, 2.4):
            r0 = ((1 - ri) / (1 + ri)) ** 2
            assert abs(O.lib().oracle_reflectance(float(cos), ri) - (r0 + (1 - r0) * (1 - cos) ** 5)) < 1e-10


# ---- RayColor / AmbientLight, ray/objects_test.go:227-320 -----------------------------------------------
def test_ray_color_properties(O):
    sc = O.rich_scene(2)
    assert O.ray_color(sc, (0, 0, 0), (0, 0, -1), 0).tolist() == [0, 0, 0]  # depth 0 -> exactly black
    empty = O.FlatScene([], [], [], [], [], np.zeros((0, 4)))
    up, down = O.ray_color(empty, (0, 0, 0), (0, 1, -1), 5), O.ray_color(empty, (0, 0, 0), (0, -1, -1), 5)
    assert up[2] == 1.0 and down[2] == 1.0 and up[0] < down[0]  # a larger up => more ColorB (blue): less red
    for i in range(200):
        c = O.ray_color(sc, (13, 2, 3), (-13 + 0.01 * i, -2, -3), 50, idx=i)
        assert np.all(c >= 0) and np.all(c <= 1)


# ---- Tracer, ray/tracer_test.go ----------------------------------------------------------------------
def test_render_lines_touches_only_its_rows(O):
    sc = O.rich_scene(2)
    cam = O.camera_init(10, 10, focal_length=5, vfov=30.0)
    p = O.make_params(10, 10, spp=1, max_depth=10, seed=5)
    img = np.zeros((10, 10, 4), dtype=np.uint8)
    O.render_lines(sc, cam, p, 0, 0, 3, img)
    assert np.all(img[:3, :, 3] == 255) and not img[3:].any()


def test_workers_and_chunking(O):
    # workers 1/2/20 incl. more workers than rows all complete (tracer_test.go:188-222); W==1 is
    # RenderLines(0,0,h); W>1 chunks are independent streams indexed by their start row (tracer.go:93-121)
    sc = O.rich_scene(2)
    w, h = 24, 10
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    imgs = {}
    for W in (1, 2, 20):
        p = O.make_params(w, h, spp=2, max_depth=8, seed=9, num_workers=W, stream_mode=0)
        imgs[W], _, st = O.render(sc, cam, p)
        assert st["paths"] == w * h * 2 and np.all(imgs[W][:, :, 3] == 255)
    one = np.zeros_like(imgs[1])
    O.render_lines(sc, cam, O.make_params(w, h, spp=2, max_depth=8, seed=9), 0, 0, h, one)
    assert np.array_equal(one, imgs[1])
    assert np.array_equal(imgs[2], imgs[20])  # both chunk at max(4, h/(4W)) = 4 rows
    chunked = np.zeros_like(one)
    for y in range(0, h, 4):
        O.render_lines(sc, cam, O.make_params(w, h, spp=2, max_depth=8, seed=9), y, y, min(y + 4, h), chunked)
    assert np.array_equal(chunked, imgs[2])


def test_per_sample_streams_are_partition_independent(O):
    sc = O.rich_scene(2)
    w, h = 32, 18
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    full, _, _ = O.render(sc, cam, O.make_params(w, h, spp=3, max_depth=20, seed=2, num_workers=1, stream_mode=1))
    parts = np.zeros_like(full)
    p = O.make_params(w, h, spp=3, max_depth=20, seed=2, stream_mode=1)
    for y0, y1 in ((0, 5), (5, 6), (6, 18)):
        O.render_lines(sc, cam, p, 123, y0, y1, parts)
    assert np.array_equal(full, parts)


def test_fma_mode_within_tolerance_of_strict(O):
    # north_star tolerance: 8-bit image within 1 LSB on >= 99.9 % of pixels
    sc = O.rich_scene(2)
    w, h = 200, 112
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    a, _, _ = O.render(sc, cam, O.make_params(w, h, spp=8, max_depth=50, seed=2, num_workers=8, stream_mode=1, fma_mode=0))
    b, _, _ = O.render(sc, cam, O.make_params(w, h, spp=8, max_depth=50, seed=2, num_workers=8, stream_mode=1, fma_mode=1))
    d = np.abs(a.astype(int) - b.astype(int)).max(axis=2)
    assert (d <= 1).mean() >= 0.999


def test_oracle_matches_committed_golden(O):
    g = np.load(os.path.join(GOLDEN, "oracle_small.npz"))
    w, h, spp, depth, seed = (int(g[k]) for k in ("width", "height", "spp", "depth", "seed"))
    sc = O.rich_scene(seed)
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    for name, mode, fma, workers in (("ps_strict", 1, 0, 4), ("ps_fma", 1, 1, 4), ("ref_w1", 0, 0, 1), ("ref_w4", 0, 0, 4)):
        img, hdr, st = O.render(sc, cam, O.make_params(w, h, spp=spp, max_depth=depth, seed=seed, num_workers=workers,
                                                       stream_mode=mode, fma_mode=fma), want_hdr=True)
        assert np.array_equal(img, g[name + "_rgba"]) and np.array_equal(hdr, g[name + "_hdr"]), name
        assert st["segments"] == int(g[name + "_segments"])
    assert np.array_equal(O.rng_u64(5, 42, 32), g["rng_u64"]) and np.array_equal(O.rng_norm(5, 42, 4096), g["rng_norm"])


# ---- "next" row 8(f)-1 restatements (third-party x/image/draw; parity unpinned, properties only) ---------------
def test_bilinear_scale_properties(O):
    c = np.full((90, 160, 4), 137, dtype=np.uint8)
    c[..., 3] = 255
    assert np.unique(O.bilinear_scale(c, 40, 30)[..., :3]).tolist() == [137]  # constants are preserved
    rng = np.random.default_rng(0)
    src = rng.integers(0, 256, (64, 96, 4), dtype=np.uint8)
    src[..., 3] = 255
    same = O.bilinear_scale(src, 96, 64)  # scale 1: the triangle filter degenerates to the identity
    assert np.array_equal(same, src)
    yy, xx = np.mgrid[0:64, 0:96]
    smooth = np.stack([xx * 2, yy * 3, (xx + yy), np.full_like(xx, 255)], axis=-1).astype(np.uint8)
    half = O.bilinear_scale(smooth, 48, 32).astype(int)
    box = smooth.reshape(32, 2, 48, 2, 4).mean(axis=(1, 3))
    assert np.abs(half[1:-1, 1:-1, :3] - box[1:-1, 1:-1, :3]).max() <= 1.0 and (half[..., 3] == 255).all()  # linear ramps survive
    frame = O.ansi_halfblocks(half.astype(np.uint8))
    assert len(frame) == 16 * (48 * 41 + 5) and frame.startswith(b"\x1b[48;2;") and frame.endswith(b"\x1b[0m\n")

"""Closes the "parity unpinned" items of DESIGN.md section 2 once a maintainer with a Go toolchain has produced
tests/golden/go_vectors.json with tools/go_vectors/main.go (run inside a fortio/tray checkout; this container has no Go).
Skipped while that file is absent. Every vector is compared with the C oracle (CPU) and, with -m gpu, with the device
generators and the reference-stream conformance kernel; for Rand.InDisc / Rand.UnitVector -- whose bodies live in
fortio.org/rand v1.1.0, outside the reference tree -- every candidate body the oracle and the device carry is tried and the
test names the one that matches, so that flipping tray_configure(TRAY_CFG_INDISC / TRAY_CFG_UNITVEC) (and the oracle's
default) is all that is left to do."""
import json
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN

PATH = os.path.join(GOLDEN, "go_vectors.json")
pytestmark = pytest.mark.skipif(not os.path.exists(PATH), reason="tests/golden/go_vectors.json absent (needs a Go toolchain: tools/go_vectors/main.go)")


@pytest.fixture(scope="module")
def V():
    return json.load(open(PATH))


def f64(h):
    return struct.unpack("<d", struct.pack("<Q", int(h, 16)))[0]


def matching_variants(O, V, kind):
    """Which oracle bodies reproduce the Go streams of `kind` ('unit_vector' or 'in_disc') bit for bit."""
    ok = []
    for v in range(3):
        O.lib().oracle_set_variants(v if kind == "in_disc" else 0, v if kind == "unit_vector" else 0)
        good = True
        for s in V["streams"]:
            want = np.array(s[kind])
            got = O.rng_unit_vectors(s["idx"], s["seed"], len(want)) if kind == "unit_vector" else O.rng_in_disc(s["idx"], s["seed"], 0.5, len(want))
            good = good and np.array_equal(got, want)
        if good:
            ok.append(v)
    O.lib().oracle_set_variants(0, 0)
    return ok


def test_oracle_streams_match_go(O, V):
    for s in V["streams"]:
        assert [int(x) for x in s["pcg_uint64"]] == [int(x) for x in O.rng_u64(s["idx"], s["seed"], len(s["pcg_uint64"]))]
        assert np.array_equal(O.rng_f64(s["idx"], s["seed"], len(s["float64"])), np.array(s["float64"])), "seeding: NewIdx(idx, seed) != NewPCG(idx, seed)"
        assert np.array_equal(O.rng_norm(s["idx"], s["seed"], len(s["pcg_norm"])), np.array(s["pcg_norm"])), "NormFloat64 (ziggurat / exp form)"


@pytest.mark.parametrize("kind", ["unit_vector", "in_disc"])
def test_oracle_wrapper_body_is_the_go_one(O, V, kind):
    ok = matching_variants(O, V, kind)
    assert ok, "no candidate body of Rand.%s reproduces Go's stream: restate it from fortio.org/rand and add it as a variant" % kind
    assert ok == [0], ("Go's Rand.%s is candidate body %s, not the default 0: make it the default in oracle/tray_oracle.c and in "
                       "tray_device.cuh (tray_configure(TRAY_CFG_%s, %d) selects it today)" % (kind, ok, "UNITVEC" if kind == "unit_vector" else "INDISC", ok[0]))


def test_oracle_srgb_exp_pow_tan_match_go(O, V):
    x = np.array([f64(h) for h in V["linear_to_srgb_inputs"]])
    assert np.array_equal(O.linear_to_srgb(x), np.array(V["linear_to_srgb"]["out"], dtype=np.uint8))
    L = O.lib()
    import ctypes
    L.oracle_go_exp.restype = ctypes.c_double
    L.oracle_go_exp.argtypes = [ctypes.c_double]
    bad = [h for h, e in V["exp"] if L.oracle_go_exp(f64(h)) != f64(e)]
    assert not bad, "math.Exp on this GOARCH (%s) differs from the pure-Go msun form at %d of %d points (assembly archExp)" % (V["goarch"], len(bad), len(V["exp"]))


def test_oracle_scene_and_render_match_go(O, V):
    for seed, n in V["rich_scene_objects"].items():
        assert O.rich_scene(int(seed)).n == n
    r = V["render"]
    w, h = r["width"], r["height"]
    ref, _, _ = O.render(O.rich_scene(2), O.camera_init(w, h, **O.RICH_CAMERA), O.make_params(w, h, spp=4, max_depth=50, seed=2, num_workers=1, stream_mode=0))
    want = np.array(r["pix"], dtype=np.uint8).reshape(h, w, 4)
    d = np.abs(ref.astype(int) - want.astype(int)).max(axis=2)
    assert (d == 0).all(), "%s: %d of %d pixels differ (max %d LSB)" % (r["args"], int((d > 0).sum()), d.size, int(d.max()))


@pytest.mark.gpu
def test_device_matches_go(ctx, O, V):
    from tray_b200 import rand, ray
    for s in V["streams"]:
        assert [int(x) for x in s["pcg_uint64"]] == [int(x) for x in ctx.rng_dump(0, s["idx"], s["seed"], len(s["pcg_uint64"]))]
        assert np.array_equal(ctx.rng_dump(2, s["idx"], s["seed"], len(s["pcg_norm"])), np.array(s["pcg_norm"]))
        assert np.array_equal(ctx.rng_dump(3, s["idx"], s["seed"], len(s["unit_vector"])), np.array(s["unit_vector"]))
        assert np.array_equal(ctx.rng_dump(4, s["idx"], s["seed"], len(s["in_disc"]), 0.5), np.array(s["in_disc"]))
    x = np.array([f64(h) for h in V["linear_to_srgb_inputs"]])
    assert np.array_equal(ctx.linear_to_srgb(x), np.array(V["linear_to_srgb"]["out"], dtype=np.uint8))
    r = V["render"]
    w, h = r["width"], r["height"]
    t = ray.New(w, h)
    t.Camera = ray.RichSceneCamera()
    t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.StreamMode, t.NumWorkers = 50, 4, 2, ray.STREAM_REFERENCE, 1
    img = t.Render(ray.RichScene(rand.New(2)))
    assert np.array_equal(img, np.array(r["pix"], dtype=np.uint8).reshape(h, w, 4)), r["args"]

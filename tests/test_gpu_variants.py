"""The alternative bodies of fortio.org/rand's Rand.InDisc / Rand.UnitVector (tray_configure(TRAY_CFG_INDISC / TRAY_CFG_UNITVEC)):
those two bodies are not in the reference tree and no reference test pins their values (call sites ray/tracer.go:138,
ray/camera.go:128, ray/rand.go:31), so the device carries every candidate the oracle knows. Each must be bit-identical to
the oracle's: generator streams, per-sample renders and the reference-stream conformance kernel. When real Go vectors
arrive (tools/go_vectors, tests/test_go_vectors.py) the matching body is selected with a flag."""
import itertools

import numpy as np
import pytest

from tray_b200 import _lib, rand, ray

pytestmark = pytest.mark.gpu


@pytest.fixture
def variants(ctx, O):
    def set_both(indisc, unitvec):
        O.lib().oracle_set_variants(indisc, unitvec)
        ctx.configure(_lib.CFG_INDISC, indisc)
        ctx.configure(_lib.CFG_UNITVEC, unitvec)
    yield set_both
    set_both(0, 0)


@pytest.mark.parametrize("indisc,unitvec", [(i, u) for i, u in itertools.product(range(3), range(3)) if (i, u) != (0, 0)])
def test_variant_streams_and_renders_match_the_oracle(ctx, O, variants, indisc, unitvec):
    variants(indisc, unitvec)
    assert ctx.query(_lib.CFG_INDISC) == indisc and ctx.query(_lib.CFG_UNITVEC) == unitvec
    for idx, seed in ((0, 2), (5, 42), (2 ** 63 + 1, 7)):
        assert np.array_equal(ctx.rng_dump(3, idx, seed, 3000), O.rng_unit_vectors(idx, seed, 3000))
        assert np.array_equal(ctx.rng_dump(4, idx, seed, 3000, 0.5), O.rng_in_disc(idx, seed, 0.5, 3000))
    w, h, spp, depth = 96, 54, 6, 30
    cam = O.camera_init(w, h, **O.RICH_CAMERA)
    for mode, workers in ((ray.STREAM_PER_SAMPLE, 8), (ray.STREAM_REFERENCE, 1), (ray.STREAM_REFERENCE, 3)):
        t = ray.New(w, h)
        t.Camera = ray.RichSceneCamera()
        t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.StreamMode, t.NumWorkers = depth, spp, 2, mode, workers
        img = t.Render(ray.RichScene(rand.New(2))).copy()
        ref, hdr, st = O.render(O.rich_scene(2), cam, O.make_params(w, h, spp=spp, max_depth=depth, seed=2, num_workers=workers,
                                                                    stream_mode=1 if mode == ray.STREAM_PER_SAMPLE else 0), want_hdr=True)
        assert np.array_equal(img, ref) and np.array_equal(ctx.read_hdr(w, h), hdr), (mode, workers)
        assert t.Stats["segments"] == st["segments"]


def test_variants_change_the_image_and_default_is_restored(ctx, O, variants):
    w, h = 64, 36

    def render():
        t = ray.New(w, h)
        t.Camera = ray.RichSceneCamera()
        t.MaxDepth, t.NumRaysPerPixel, t.Seed = 12, 4, 2
        return t.Render(ray.RichScene(rand.New(2))).copy()
    base = render()
    variants(1, 0)
    a = render()
    variants(0, 2)
    b = render()
    variants(0, 0)
    assert not np.array_equal(a, base) and not np.array_equal(b, base) and np.array_equal(render(), base)
    with pytest.raises(ray.TrayError):
        ctx.configure(_lib.CFG_INDISC, 3)

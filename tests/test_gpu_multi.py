"""Multi-device context (one process, G devices): tile mode (interleaved rows, device-to-host gather only) and
sample-split mode (partial sums reduced over NVLink peer pointers inside the combine kernel). Needs >= 2 GPUs."""
import numpy as np
import pytest

from tray_b200 import rand, ray

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _tracer(ctx, w, h, spp, depth, split=ray.SPLIT_TILES, precision=ray.FP64_STRICT):
    t = ray.New(w, h)
    t.Camera = ray.RichSceneCamera()
    t.MaxDepth, t.NumRaysPerPixel, t.Seed, t.Precision, t.SplitMode, t.Context = depth, spp, 2, precision, split, ctx
    return t


@pytest.mark.parametrize("g", [2, 4, 8])
def test_tile_mode_is_bit_identical_for_any_device_count(ctx, g):
    if _ngpu() < g:
        pytest.skip("needs %d GPUs" % g)
    scene = ray.RichScene(rand.New(2))
    w, h, spp, depth = 320, 181, 8, 50
    multi = ray.Context(list(range(g)))
    t = _tracer(multi, w, h, spp, depth)
    img = t.Render(scene).copy()
    one = _tracer(ctx, w, h, spp, depth).Render(scene).copy()
    assert t.Stats["n_devices"] == g and np.array_equal(img, one)
    one_hdr = ctx.read_hdr(w, h)   # (ctx rendered `one` last: its linear-HDR means are still there)
    assert np.array_equal(multi.read_hdr(w, h), one_hdr)
    # reference-stream mode on a multi-device context runs on device 0 only and must still return cleanly
    t.StreamMode, t.NumWorkers = ray.STREAM_REFERENCE, 3
    ref_multi = t.Render(scene).copy()
    r1 = _tracer(ctx, w, h, spp, depth)
    r1.StreamMode, r1.NumWorkers = ray.STREAM_REFERENCE, 3
    assert np.array_equal(ref_multi, r1.Render(scene))
    multi.close()


@pytest.mark.parametrize("g", [2, 8])
def test_sample_split_matches_within_rounding(ctx, g):
    if _ngpu() < g:
        pytest.skip("needs %d GPUs" % g)
    scene = ray.RichScene(rand.New(2))
    w, h, spp, depth = 320, 181, 16, 50
    ref_t = _tracer(ctx, w, h, spp, depth)
    one = ref_t.Render(scene).copy()
    hdr_one = ctx.read_hdr(w, h)
    multi = ray.Context(list(range(g)))
    t = _tracer(multi, w, h, spp, depth, split=ray.SPLIT_SAMPLES)
    img = t.Render(scene).copy()
    hdr = multi.read_hdr(w, h)
    # same samples, summed per device then across devices: only the summation order differs (a few ulp)
    assert t.Stats["segments"] == ref_t.Stats["segments"]
    assert np.max(np.abs(hdr - hdr_one) / np.maximum(np.abs(hdr_one), 1e-300)) < 1e-13
    d = np.abs(img.astype(int) - one.astype(int)).max(axis=2)
    assert (d <= 1).all() and (d == 0).mean() > 0.999
    multi.close()


def test_png_of_a_tile_split_frame_gathers_on_device_zero(ctx):
    """-save with the frame spread over devices: row bands are gathered onto device 0 by peer copies, then encoded there."""
    import io
    from PIL import Image
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    scene = ray.RichScene(rand.New(2))
    w, h, spp, depth = 320, 181, 4, 20
    multi = ray.Context([0, 1])
    t = _tracer(multi, w, h, spp, depth)
    img = t.Render(scene).copy()
    data, _ = multi.encode_png(w, h)
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")), img[:, :, :3])
    multi.close()


def test_progress_counts_every_device(ctx):
    """tray_progress on a multi-device context sums the per-device counters (each device has its own side stream and
    pinned word); after the render it equals the pixel count (ProgressFunc deltas sum to w*h, ray/tracer_test.go:172-186)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    scene = ray.RichScene(rand.New(2))
    w, h = 320, 181
    multi = ray.Context([0, 1])
    t = _tracer(multi, w, h, 8, 50)
    seen = []
    t.ProgressFunc = seen.append
    t.Render(scene)
    assert sum(seen) == w * h and multi.progress() == w * h
    multi.close()

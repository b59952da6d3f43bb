"""The -save path on the device (tray_encode_png <- SaveImage / png.Encode, main.go:26-36): format parity = a PNG decoder
returns exactly the rendered pixels. Checked with two independent decoders: Pillow, and a strict parser written here
(signature, chunk lengths and CRCs via zlib.crc32, zlib.decompress -- which verifies the Adler-32 -- and the five PNG
filters undone by hand)."""
import io
import struct
import zlib

import numpy as np
import pytest

from tray_b200 import rand, ray
from test_gpu_parity import tracer

pytestmark = pytest.mark.gpu


def parse_png(data):
    """Strict decoder for 8-bit truecolour, non-interlaced PNG. Returns (h, w, 3) uint8 and the list of chunk names."""
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks, idat, ihdr = 8, [], b"", None
    while pos < len(data):
        n, = struct.unpack(">I", data[pos:pos + 4])
        name = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(name + body) == crc, "bad CRC in %r" % name
        chunks.append(name)
        if name == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif name == b"IDAT":
            idat += body
        pos += 12 + n
    assert pos == len(data) and chunks[0] == b"IHDR" and chunks[-1] == b"IEND"
    w, h, depth, ctype, comp, filt, interlace = ihdr
    assert (depth, ctype, comp, filt, interlace) == (8, 2, 0, 0, 0)
    raw = zlib.decompress(idat)  # raises on a bad deflate stream or Adler-32
    stride = 1 + 3 * w
    assert len(raw) == h * stride
    img = np.zeros((h, w * 3), dtype=np.int32)
    ftypes = []
    for y in range(h):
        f = raw[y * stride]
        ftypes.append(f)
        line = np.frombuffer(raw, dtype=np.uint8, count=3 * w, offset=y * stride + 1).astype(np.int32)
        up = img[y - 1] if y > 0 else np.zeros(3 * w, dtype=np.int32)
        if f == 0:
            img[y] = line
        elif f == 2:
            img[y] = (line + up) & 255
        else:
            cur = img[y]
            for i in range(3 * w):
                a = cur[i - 3] if i >= 3 else 0
                b = up[i]
                c = up[i - 3] if i >= 3 else 0
                if f == 1:
                    pred = a
                elif f == 3:
                    pred = (a + b) >> 1
                else:
                    assert f == 4
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
    return img.reshape(h, w, 3).astype(np.uint8), chunks, ftypes, len(raw)


@pytest.mark.parametrize("w,h,spp", [(1, 1, 1), (3, 2, 2), (64, 36, 4), (333, 17, 2), (257, 129, 1)])
def test_png_decodes_to_the_rendered_pixels(ctx, w, h, spp):
    from PIL import Image
    t = tracer(w, h, spp, 12)
    img = t.Render(ray.RichScene(rand.New(2))).copy()
    data, ms = ctx.encode_png(w, h)
    got, chunks, ftypes, raw_len = parse_png(data)
    assert chunks == [b"IHDR", b"IDAT", b"IEND"]
    assert np.array_equal(got, img[:, :, :3])
    pil = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(pil, img[:, :, :3])
    assert set(ftypes) <= {0, 1, 2, 3, 4}


def test_png_of_a_smooth_frame_compresses_and_save_image_writes_it(ctx, tmp_path):
    """Sky only (empty scene): gradients -> the filters + entropy code must beat the raw size by a wide margin."""
    from PIL import Image
    w, h = 640, 360
    t = tracer(w, h, 1, 5)
    img = t.Render(ray.Scene()).copy()
    fname = str(tmp_path / "sky.png")
    ms = ray.SaveImage(t, fname)
    data = open(fname, "rb").read()
    assert len(data) < 0.35 * w * h * 3 and ms > 0
    assert np.array_equal(np.asarray(Image.open(fname).convert("RGB")), img[:, :, :3])
    got, _, ftypes, _ = parse_png(data[:])
    assert np.array_equal(got, img[:, :, :3])


def test_png_full_size_frame(ctx):
    """config 2 geometry (1920x1080, a few rays/pixel): many deflate blocks, multi-MB stream, CRC combine over ~1500 pieces."""
    from PIL import Image
    w, h = 1920, 1080
    t = tracer(w, h, 2, 12)
    img = t.Render(ray.RichScene(rand.New(2))).copy()
    data, ms = ctx.encode_png(w, h)
    pil = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert np.array_equal(pil, img[:, :, :3])
    ref = len(zlib.compress(np.ascontiguousarray(img[:, :, :3]).tobytes(), 6))
    print("png %d bytes in %.3f ms on device; zlib-6 of the unfiltered pixels: %d bytes" % (len(data), ms, ref))
    assert len(data) < 1.05 * w * h * 3


def test_png_error_paths(ctx):
    import ctypes as C
    t = tracer(32, 16, 1, 5)
    t.RenderLines  # noqa: B018
    t.Render(ray.RichScene(rand.New(2)))
    n = C.c_size_t()
    small = np.zeros(16, dtype=np.uint8)
    rc = ctx._L.tray_encode_png(ctx.handle, small.ctypes.data_as(C.c_void_p), 16, C.byref(n), None)
    assert rc != 0 and n.value > 16                      # too small: the size needed is reported
    rc = ctx._L.tray_encode_png(ctx.handle, None, 0, C.byref(n), None)
    assert rc == 0 and 60 < n.value <= ctx._L.tray_png_bound(32, 16)
    p = t._params(4, 12)                                 # a partial frame cannot be saved
    ctx.render(t.to_c(), p, None)
    with pytest.raises(ray.TrayError):
        ctx.encode_png(32, 16)


def _fib_image(w, h, n_sym):
    """Mostly constant frame whose byte histogram follows Fibonacci counts for n_sym values: the Huffman tree of a block is
    ~n_sym levels deep before the 15-bit limit is applied (the case tests/test_png_huffman_model.py checks on the CPU)."""
    img = np.zeros((h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    flat = img[:, :, :3].reshape(-1)
    fib, pos = [1, 1], 5
    while len(fib) < n_sym:
        fib.append(fib[-1] + fib[-2])
    for k, cnt in enumerate(fib):
        # isolated bytes (every 7th) so that the Sub/Up/Paeth residuals keep the value and its negative rare as well
        for _ in range(min(cnt, 4000)):
            if pos >= len(flat):
                break
            flat[pos] = 3 + 5 * k
            pos += 7
    return img


@pytest.mark.parametrize("n_sym,w,h", [(22, 640, 64), (30, 1024, 128), (12, 64, 8)])
def test_png_length_limited_codes_on_skewed_histograms(ctx, n_sym, w, h):
    from PIL import Image
    img = _fib_image(w, h, n_sym)
    ctx.upload_frame(img)
    data, _ = ctx.encode_png(w, h)
    got, chunks, _, _ = parse_png(data)                     # zlib.decompress raises on an over-subscribed / incomplete code
    assert np.array_equal(got, img[:, :, :3])
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")), img[:, :, :3])


def test_png_of_uploaded_random_and_flat_frames(ctx):
    rs = np.random.RandomState(3)
    for img in (rs.randint(0, 256, (37, 53, 4)).astype(np.uint8), np.full((16, 16, 4), 200, dtype=np.uint8), np.zeros((5, 300, 4), dtype=np.uint8)):
        img[..., 3] = 255
        ctx.upload_frame(img)
        h, w = img.shape[:2]
        data, _ = ctx.encode_png(w, h)
        got, _, _, _ = parse_png(data)
        assert np.array_equal(got, img[:, :, :3])
        back = np.zeros_like(img)
        ctx.read_image(back)
        assert np.array_equal(back, img)

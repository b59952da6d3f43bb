"""The strict PNG parser that judges the device encoder (tests/test_gpu_png.py::parse_png) is itself checked here, on the
CPU, against files written by Pillow (which uses all five filter types on natural images) and against corrupted files."""
import io
import struct
import zlib

import numpy as np
import pytest

from test_gpu_png import parse_png


def _pillow_png(img):
    from PIL import Image
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="PNG")
    return buf.getvalue()


def _smooth_image(w, h, seed):
    rs = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    base = np.stack([128 + 100 * np.sin(x / 9.0 + seed), 128 + 100 * np.cos(y / 7.0), (x * 3 + y * 5) % 256], axis=-1)
    return np.clip(base + rs.normal(0, 3, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("w,h,seed", [(1, 1, 0), (3, 2, 1), (97, 41, 2), (256, 64, 3)])
def test_parser_decodes_pillow_files_exactly(w, h, seed):
    img = _smooth_image(w, h, seed)
    got, chunks, ftypes, raw_len = parse_png(_pillow_png(img))
    assert np.array_equal(got, img) and chunks[0] == b"IHDR" and chunks[-1] == b"IEND" and raw_len == h * (1 + 3 * w)


def test_parser_sees_every_filter_type():
    """Hand-built file with one scanline per filter type (0..4): the parser must undo each of them."""
    w = 16
    rs = np.random.RandomState(7)
    img = rs.randint(0, 256, (5, w, 3)).astype(np.uint8)
    raw = bytearray()
    prev = np.zeros(3 * w, dtype=np.int32)
    for f in range(5):
        cur = img[f].reshape(-1).astype(np.int32)
        line = np.zeros(3 * w, dtype=np.int32)
        for i in range(3 * w):
            a = cur[i - 3] if i >= 3 else 0
            b = prev[i]
            c = prev[i - 3] if i >= 3 else 0
            if f == 0:
                pred = 0
            elif f == 1:
                pred = a
            elif f == 2:
                pred = b
            elif f == 3:
                pred = (a + b) >> 1
            else:
                p = a + b - c
                pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
            line[i] = (cur[i] - pred) & 255
        raw += bytes([f]) + bytes(line.astype(np.uint8))
        prev = cur

    def chunk(name, body):
        return struct.pack(">I", len(body)) + name + body + struct.pack(">I", zlib.crc32(name + body))
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, 5, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b"")
    got, _, ftypes, _ = parse_png(data)
    assert ftypes == [0, 1, 2, 3, 4] and np.array_equal(got, img)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(io.BytesIO(data)).convert("RGB")), img)   # and Pillow reads the same pixels


def test_parser_rejects_corruption():
    data = bytearray(_pillow_png(_smooth_image(40, 20, 5)))
    bad_crc = bytearray(data)
    bad_crc[-5] ^= 1                       # IEND CRC
    with pytest.raises(AssertionError):
        parse_png(bytes(bad_crc))
    bad_body = bytearray(data)
    bad_body[60] ^= 0x40                   # a byte inside IDAT: chunk CRC no longer matches
    with pytest.raises(AssertionError):
        parse_png(bytes(bad_body))

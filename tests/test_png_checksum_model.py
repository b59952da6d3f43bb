"""Python model of the device's checksum combination (tray_b200/csrc/tray_png.cuh): CRC-32 of the IDAT chunk from raw CRCs
of 4 kB pieces folded with x^(8n) mod P shifts by 256 threads and then by one, and Adler-32 from per-row partial sums.
Checked against zlib for the sizes where the folding has edges (one piece, exact multiples, fewer pieces than threads)."""
import random
import zlib

POLY = 0xEDB88320
CHUNK = 4096
M32 = 0xFFFFFFFF


def mulmod(a, b):
    p = 0
    for _ in range(32):
        if a & 0x80000000:
            p ^= b
        a = (a << 1) & M32
        b = (b >> 1) ^ POLY if b & 1 else b >> 1
    return p


def xpow8n(n):
    r, sq = 0x80000000, 0x00800000
    while n:
        if n & 1:
            r = mulmod(r, sq)
        sq = mulmod(sq, sq)
        n >>= 1
    return r


TABLE = []
for i in range(256):
    c = i
    for _ in range(8):
        c = POLY ^ (c >> 1) if c & 1 else c >> 1
    TABLE.append(c)


def raw_crc(data):   # init 0, no final xor (png_crc_kernel)
    c = 0
    for b in data:
        c = TABLE[(c ^ b) & 255] ^ (c >> 8)
    return c


def device_crc(data):
    total = len(data)
    n_pieces = (total + CHUNK - 1) // CHUNK
    piece = [raw_crc(data[k * CHUNK:(k + 1) * CHUNK]) for k in range(n_pieces)]
    per = (n_pieces + 255) // 256
    xfull = xpow8n(CHUNK)
    s_crc, s_mul = [], []
    for tid in range(256):          # png_finish_kernel, mode 1
        acc, ln = 0, 0
        for k in range(tid * per, min((tid + 1) * per, n_pieces)):
            plen = CHUNK if (k + 1) * CHUNK <= total else total - k * CHUNK
            acc = mulmod(acc, xfull if plen == CHUNK else xpow8n(plen)) ^ piece[k]
            ln += plen
        s_crc.append(acc)
        s_mul.append(xpow8n(ln))
    c = 0
    for t in range(256):
        c = mulmod(c, s_mul[t]) ^ s_crc[t]
    return c ^ mulmod(0xFFFFFFFF, xpow8n(total)) ^ 0xFFFFFFFF


def test_crc_combination_matches_zlib():
    rnd = random.Random(5)
    for n in (1, 4, 4095, 4096, 4097, 8192, 3 * 4096 + 17, 255 * 4096, 256 * 4096, 256 * 4096 + 1, 257 * 4096 + 4000, 700 * 4096 + 5):
        data = bytes(rnd.getrandbits(8) for _ in range(min(n, 70000))) * (n // min(n, 70000)) + bytes(rnd.getrandbits(8) for _ in range(n % min(n, 70000)))
        assert len(data) == n
        assert device_crc(data) == zlib.crc32(data), n


def test_adler_row_combination_matches_zlib():
    rnd = random.Random(6)
    for L, h in ((4, 1), (10, 3), (1 + 3 * 1920, 7), (65521, 2), (70000, 3)):
        rows = [bytes(rnd.getrandbits(8) for _ in range(L)) for _ in range(h)]
        A, B, M = 1, 0, 65521
        for r in rows:              # png_filter_kernel partial sums + png_layout_kernel combination
            s1 = sum(r) % M
            s2 = sum((L - i) * b for i, b in enumerate(r)) % M
            B = (B + (L % M) * A + s2) % M
            A = (A + s1) % M
        assert ((B << 16) | A) == zlib.adler32(b"".join(rows)), (L, h)

"""CPU checks of the two-level cluster closest hit (cluster_scan in tray_b200/csrc/tray_kernels.cuh, tables built by
build_clusters in tray_api.cu), the structure that replaces the linear scan of Scene.Hit (ray/objects.go:37-46).

1. The tables the library stages (read through the device-free tray_cluster_tables entry point) partition the scene:
   every sphere sits in exactly one slot, empty slots point at the never-hit padding entry, chunk boxes contain their
   spheres, group boxes their chunks, spheres the fp32 filter cannot bound sit in always-groups.
2. The conservative fp32 slab test is replayed with exact single rounding (fused multiply-adds through rational arithmetic,
   helpers shared with test_prefilter_model.py) and attacked with grazing rays, origins on and inside surfaces, axis-parallel
   and near-axis-parallel directions, far origins: whenever the strict fp64 Sphere.Hit of the reference (oracle/pyref.py =
   Go/amd64 semantics) reports a hit, neither the sphere's chunk box nor its group box may be called "missed".
No compute call reaches a GPU here."""
import ctypes as C
import math

import numpy as np
import pytest

from test_prefilter_model import F32, U32, fma32, sign

from tray_b200 import _lib


def cluster_tables(cx, cy, cz, r):
    L = _lib.lib()
    d = _lib.SceneDesc()
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (cx, cy, cz, r)]
    d.n = len(arrs[0])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    d.cx, d.cy, d.cz, d.radius = (p(a) for a in arrs)
    meta = (C.c_int32 * 8)()
    rr = C.c_float()
    assert L.tray_cluster_tables(C.byref(d), None, 0, meta, C.byref(rr)) == 0
    blob = np.zeros(meta[0] * 4, dtype=np.float32)
    assert L.tray_cluster_tables(C.byref(d), p(blob), blob.size, meta, C.byref(rr)) == 0
    m = dict(zip(("blob_f4", "off_box2", "off_box1", "off_ids", "real_groups", "always_groups", "always_last", "unfilterable"), list(meta)))
    n_chunks = m["off_box2"] // 8
    f4 = blob.reshape(-1, 4)
    pairs = f4[:m["off_box2"]].reshape(n_chunks, 4, 2, 4)           # [chunk][pair][float4 #][lane]
    ids = blob[m["off_ids"] * 4:].view(np.uint16)[:n_chunks * 8].reshape(n_chunks, 8).astype(int)

    def boxes(off, count):
        b = f4[off:off + count // 2 * 3].reshape(count // 2, 3, 4)
        c = np.zeros((count, 3), dtype=np.float32); e = np.zeros((count, 3), dtype=np.float32)
        for j in range(2):
            c[j::2, 0], c[j::2, 1], c[j::2, 2] = b[:, 0, j], b[:, 0, 2 + j], b[:, 1, j]
            e[j::2, 0], e[j::2, 1], e[j::2, 2] = b[:, 1, 2 + j], b[:, 2, j], b[:, 2, 2 + j]
        return c, e
    words8 = (m["real_groups"] // 8 + 7) // 8 * 8  # [box0]: one box per word of 8 groups, padded to a multiple of 8 words, right after [box1]
    off_box0 = m["off_box1"] + m["real_groups"] // 2 * 3
    assert off_box0 + words8 // 2 * 3 == m["off_ids"]
    m.update(n_chunks=n_chunks, pairs=pairs, ids=ids, box2=boxes(m["off_box2"], n_chunks), box1=boxes(m["off_box1"], m["real_groups"]),
             box0=boxes(off_box0, words8), words8=words8, r=float(rr.value))
    return m


def scenes():
    from oracle import oracle as O
    O.build()
    rs = np.random.RandomState(5)
    s = O.rich_scene(2)
    yield "rich", s.cx, s.cy, s.cz, s.r
    yield "empty", [], [], [], []
    yield "one", [0.5], [1.0], [-2.0], [0.25]
    for n in (7, 8, 9, 63, 64, 65, 200, 513, 700, 4100, 10001):
        yield "rand%d" % n, rs.uniform(-20, 20, n), rs.uniform(0, 3, n), rs.uniform(-20, 20, n), rs.choice([0.2, 0.3, 1.5, -0.4], n)
    n = 40  # many unfilterable spheres (far away / huge) plus many large ones
    yield "far", np.r_[rs.uniform(-20, 20, n), rs.uniform(300, 900, 20)], rs.uniform(-5, 5, n + 20), rs.uniform(-20, 20, n + 20), \
        np.r_[rs.choice([0.2, 5.0, 30.0], n), rs.uniform(0.1, 2, 20)]
    yield "nan", [0.0, float("nan"), 1.0], [0.0, 0.0, float("inf")], [0.0, 1.0, 2.0], [1.0, 1.0, 1.0]


@pytest.mark.parametrize("name,cx,cy,cz,r", list(scenes()), ids=lambda v: v if isinstance(v, str) else None)
def test_cluster_tables_partition_the_scene(name, cx, cy, cz, r):
    cx, cy, cz, r = (np.asarray(a, dtype=np.float64) for a in (cx, cy, cz, r))
    n = len(cx)
    m = cluster_tables(cx, cy, cz, r)
    ids, pairs = m["ids"], m["pairs"]
    assert m["real_groups"] % 8 == 0 and m["n_chunks"] == 8 * (m["real_groups"] + m["always_groups"]) or m["n_chunks"] % 8 == 0
    real = ids[ids < n]
    assert sorted(real.tolist()) == list(range(n)), "every sphere in exactly one slot"
    assert (ids[ids >= n] == n).all(), "empty slots point at the padding entry"
    with np.errstate(invalid="ignore"):
        filt = (np.maximum(np.abs(cx), np.maximum(np.abs(cy), np.abs(cz))) <= 256) & (r * r <= 256)
    assert m["unfilterable"] == int((~filt).sum())
    c2, e2 = m["box2"]
    c1, e1 = m["box1"]
    c0, e0 = m["box0"]   # word boxes (the level cluster_scan_big tests first)
    first_always = m["real_groups"] * 8
    for ch in range(m["n_chunks"]):
        for u in range(8):
            i = ids[ch, u]
            p, k = divmod(u, 2)
            g0, g1 = pairs[ch, p, 0], pairs[ch, p, 1]
            entry = (g0[k], g0[2 + k], g1[k], g1[2 + k])   # cx, cy, cz, -K
            if i >= n:
                assert entry[3] == -np.inf and entry[:3] == (0, 0, 0)
                continue
            if not filt[i]:
                assert ch >= first_always and all(np.isnan(v) for v in entry)
                continue
            assert entry[:3] == (F32(cx[i]), F32(cy[i]), F32(cz[i]))
            assert entry[3] == -F32((cx[i] * cx[i] + cy[i] * cy[i] + cz[i] * cz[i]) - r[i] * r[i])
            lo = np.array([cx[i], cy[i], cz[i]]) - abs(r[i])
            hi = np.array([cx[i], cy[i], cz[i]]) + abs(r[i])
            for cc, ee in ((c2[ch], e2[ch]),) + (((c1[ch // 8], e1[ch // 8]), (c0[ch // 64], e0[ch // 64])) if ch < first_always else ()):
                assert (cc.astype(np.float64) - ee.astype(np.float64) < lo).all() and (cc.astype(np.float64) + ee.astype(np.float64) > hi).all()
                if np.isfinite(ee).all():   # (an always-chunk holding an unbounded sphere has the box "everything")
                    assert (np.abs(cc.astype(np.float64)) + ee.astype(np.float64) <= m["r"]).all()
        if (ids[ch] >= n).all():
            assert (e2[ch] == -np.inf).all(), "an empty chunk can never be hit"
    used_groups = {ch // 8 for ch in range(first_always) if (ids[ch] < n).any()}
    for g in range(m["real_groups"]):
        if g not in used_groups:
            assert (e1[g] == -np.inf).all()
    for wd in range(m["words8"]):
        if wd >= m["real_groups"] // 8 or not any(g in used_groups for g in range(8 * wd, 8 * wd + 8)):
            assert (e0[wd] == -np.inf).all(), "an empty or padding word can never be hit"
    if m["always_groups"]:
        last = [ch for ch in range(first_always + 8 * (m["always_groups"] - 1), m["n_chunks"]) if (ids[ch] < n).any()]
        assert m["always_last"] == sum(0x80 >> (ch % 8) for ch in last)


def test_rich_scene_layout():
    """The benchmark scene: 481 small spheres in 61 chunks = 8 groups behind boxes, the ground (outside the filter's range)
    and the three unit spheres in one always-chunk."""
    from oracle import oracle as O
    O.build()
    s = O.rich_scene(2)
    m = cluster_tables(s.cx, s.cy, s.cz, s.r)
    assert (m["real_groups"], m["always_groups"], m["always_last"], m["unfilterable"]) == (8, 1, 0x80, 1)
    assert sorted(m["ids"][64][m["ids"][64] < s.n].tolist()) == [0, s.n - 3, s.n - 2, s.n - 1]
    _, e2 = m["box2"]
    used = [ch for ch in range(64) if (m["ids"][ch] < s.n).any()]
    assert len(used) == 61 and float(e2[used].max()) < 4.0   # compact boxes: a chunk spans a few grid cells


# ---- exact model of the slab test ---------------------------------------------------------------------------------

def box_ray_constants(o, d, cl_r, rcp_ulps=(0, 0, 0)):
    """rcp_ulps: the device takes 1/d_k with MUFU.RCP (within 1 ulp of the quotient); the model pushes the correctly rounded
    quotient by -1, 0 or +1 ulp per axis so that the attack covers the whole error interval."""
    a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    ln = math.sqrt(a)                                   # Unit(D) of the RayColor step (ray/vec3.go:117-120), fp64
    fd = [F32(d[k] / ln) for k in range(3)]
    lim = F32(8.6736174e-19)
    inv = []
    for k in range(3):
        v = fd[k]
        if abs(v) < lim:
            v = F32(math.copysign(float(lim), float(v)))
        q = F32(1.0) / v
        if rcp_ulps[k]:
            q = np.nextafter(q, F32(np.inf) if rcp_ulps[k] > 0 else F32(-np.inf))
        inv.append(F32(q))
    nq = [-F32(o[k] * float(inv[k])) for k in range(3)]
    mo = F32(1.0000002) * max(abs(F32(o[0])), abs(F32(o[1])), abs(F32(o[2])))
    R = F32(cl_r) + mo
    ks = F32(16.0) * U32 * R
    sl = [ks * abs(inv[k]) for k in range(3)]
    off = not (mo < F32(1e6)) or not (0.0 < a < 1.7976931348623157e308)
    return inv, nq, sl, off


def box_says_missed(consts, c, e):
    inv, nq, sl, off = consts
    if off:
        return False
    near, far = [], []
    for k in range(3):
        A = fma32(c[k], inv[k], nq[k])
        B = fma32(e[k], abs(inv[k]), sl[k])
        near.append(A - B)
        far.append(A + B)
    tn, tf = max(near), min(far)    # no NaN for a finite ray and a finite box
    return sign(tf - tn) or sign(tf)


def test_box_test_never_culls_a_sphere_the_strict_test_would_hit():
    from oracle import pyref
    rs = np.random.RandomState(21)
    culled = kept = hits = hits_kept = 0
    for trial in range(12000):
        scale = float(rs.choice([1.0, 12.0, 200.0]))
        r = float(rs.choice([0.2, 1.0, 0.05, 0.001, 3.0]))
        c = tuple(float(v) for v in rs.uniform(-scale, scale, 3))
        m = cluster_tables([c[0]], [c[1]], [c[2]], [r])   # one sphere = one chunk: the tightest box the builder makes
        bc, be = m["box2"][0][0], m["box2"][1][0]
        cl_r = m["r"] + float(rs.choice([0.0, scale]))
        kind = trial % 6
        dirn = rs.normal(0, 1, 3)
        dirn /= np.linalg.norm(dirn)
        if kind == 0:      # grazing
            perp = np.cross(dirn, rs.normal(0, 1, 3))
            perp /= np.linalg.norm(perp)
            eps = float(rs.choice([0.0, 1e-15, 1e-12, 1e-9, 1e-7, 1e-5, 1e-3])) * float(rs.choice([-1, 1]))
            o = np.array(c) + perp * r * (1 + eps) - dirn * float(rs.uniform(0.1, 3.0) * scale)
        elif kind == 1:    # origin on the surface
            nn = rs.normal(0, 1, 3)
            nn /= np.linalg.norm(nn)
            o = np.array(c) + nn * r * (1 + float(rs.choice([0, 1e-15, -1e-15, 1e-9, -1e-9])))
            dirn = nn * float(rs.choice([1, -1])) + rs.normal(0, 1, 3) * float(rs.choice([0.01, 1.0]))
        elif kind == 2:    # origin inside
            o = np.array(c) + rs.uniform(-0.5, 0.5, 3) * r
        elif kind == 3:    # axis-parallel and nearly axis-parallel directions through / past the sphere
            ax = rs.randint(3)
            dirn = np.zeros(3)
            dirn[ax] = float(rs.choice([1, -1]))
            dirn = dirn + rs.normal(0, 1, 3) * float(rs.choice([0.0, 0.0, 1e-30, 1e-20, 1e-12, 1e-7]))
            off_ = rs.uniform(-1, 1, 3) * r * float(rs.choice([0.5, 1.0, 1.000001]))
            off_[ax] = -float(rs.choice([2.0, 50.0])) * r * dirn[ax]
            o = np.array(c) + off_
        elif kind == 4:    # far origins (large t, large |O|)
            o = np.array(c) - dirn * float(rs.choice([1e3, 1e5, 9e5])) + rs.normal(0, 1, 3) * r * 0.7
        else:
            o = rs.uniform(-1.5, 1.5, 3) * scale
        if not np.any(dirn):
            continue
        d = tuple(float(v) for v in dirn * float(rs.choice([1e-3, 1.0, 10.0])))
        o = tuple(float(v) for v in o)
        hit = pyref.sphere_hit(c, r, o, d, 1e-6, math.inf)
        hits += hit is not None
        missed = box_says_missed(box_ray_constants(o, d, cl_r, tuple(int(v) for v in rs.randint(-1, 2, 3))), bc, be)
        if missed:
            culled += 1
            assert hit is None, (trial, kind, o, d, c, r)
        else:
            kept += 1
            hits_kept += hit is not None
    assert culled > 1500 and hits > 3000 and kept > 3000, (culled, hits, kept)   # the attack reached both sides of the test


def test_box_test_on_the_benchmark_scene_keeps_every_hit_and_culls_most_chunks():
    """Rays of the benchmark scene (camera rays and rays leaving surfaces): every sphere the strict test hits lies in a
    chunk and a group the model keeps, and the boxes do their job (a ray keeps a handful of the 61 chunks)."""
    from oracle import oracle as O
    from oracle import pyref
    O.build()
    s = O.rich_scene(2)
    m = cluster_tables(s.cx, s.cy, s.cz, s.r)
    c2, e2 = m["box2"]
    c1, e1 = m["box1"]
    where = {}
    for ch in range(m["n_chunks"]):
        for i in m["ids"][ch]:
            if i < s.n:
                where[int(i)] = ch
    rs = np.random.RandomState(3)
    spheres = [((s.cx[i], s.cy[i], s.cz[i]), s.r[i]) for i in range(s.n)]
    kept_chunks = rays = 0
    for trial in range(160):
        if trial % 2 == 0:
            o = (13.0 + rs.normal() * 0.05, 2.0 + rs.normal() * 0.05, 3.0 + rs.normal() * 0.05)
            tgt = (rs.uniform(-8, 8), rs.uniform(0, 1), rs.uniform(-8, 8))
            d = tuple(tgt[k] - o[k] for k in range(3))
        else:
            i = rs.randint(1, s.n)
            nn = rs.normal(0, 1, 3); nn /= np.linalg.norm(nn)
            o = tuple(float(v) for v in np.array(spheres[i][0]) + nn * spheres[i][1])
            d = tuple(float(v) for v in nn + rs.normal(0, 1, 3) * 0.7)
        consts = box_ray_constants(o, d, m["r"])
        miss2 = [box_says_missed(consts, c2[ch], e2[ch]) for ch in range(64)]
        miss1 = [box_says_missed(consts, c1[g], e1[g]) for g in range(8)]
        rays += 1
        kept_chunks += sum(1 for ch in range(64) if not miss2[ch] and not miss1[ch // 8])
        for i in range(1, s.n - 3):
            if pyref.sphere_hit(spheres[i][0], spheres[i][1], o, d, 1e-6, math.inf) is not None:
                ch = where[i]
                assert not miss2[ch] and not miss1[ch // 8], (trial, i, ch)
    assert kept_chunks / rays < 8.0, kept_chunks / rays

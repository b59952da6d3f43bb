"""Host logic of the drop-in boundary (no GPU): the Python mirror of the Go package API, the marshalling
into the C ABI, and the shared library's exported surface. Tables follow ray/tracer_test.go and
ray/camera_test.go of the reference."""
import ctypes
import math
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tray_b200 import _lib, rand, ray

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_declared_symbol():
    so = _lib.build_library()
    L = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "tray_cuda.h")).read()
    declared = re.findall(r"TRAY_API\s+[\w\s\*]+?\b(tray_\w+)\s*\(", hdr)
    assert len(declared) >= 13 and set(declared) == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(L, name) is not None
    L.tray_abi_version.restype = ctypes.c_int
    assert L.tray_abi_version() == int(re.search(r'#define TRAY_ABI_VERSION (\d+)', hdr).group(1)) >= 2


def test_ctypes_structs_match_the_c_header():
    """sizeof/offsetof of the ABI structs as gcc sees the header == the ctypes mirrors."""
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "tray_cuda.h"
int main(void) {
  printf("%zu %zu %zu %zu\n", sizeof(tray_scene_desc), sizeof(tray_camera), sizeof(tray_params), sizeof(tray_stats));
  printf("%zu %zu %zu %zu %zu\n", offsetof(tray_params, seed), offsetof(tray_params, stream_idx), offsetof(tray_params, precision),
         offsetof(tray_params, shard_index), offsetof(tray_scene_desc, bg_a));
  printf("%zu %zu\n", offsetof(tray_params, sample_offset), offsetof(tray_params, sums_mode));
  printf("%zu %zu %zu\n", offsetof(tray_stats, kernel_ms), offsetof(tray_stats, trace_kernel_ms), offsetof(tray_stats, trace_launches));
  return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "a.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "a.out")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [int(x) for x in out]
    P, S, St = _lib.Params, _lib.SceneDesc, _lib.Stats
    want = [ctypes.sizeof(S), ctypes.sizeof(_lib.CameraC), ctypes.sizeof(P), ctypes.sizeof(St),
            P.seed.offset, P.stream_idx.offset, P.precision.offset, P.shard_index.offset, S.bg_a.offset,
            P.sample_offset.offset, P.sums_mode.offset,
            St.kernel_ms.offset, St.trace_kernel_ms.offset, St.trace_launches.offset]
    assert got == want


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ray.TrayError) as e:
        ray.Context()
    assert e.value.code == _lib.E_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tray_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".go")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("host oracle and device", ""), os.path.join(dirpath, f)


# ---- ray.New / Tracer defaults, ray/tracer_test.go:9-45,108-170 -------------------------------------
def test_new_tracer():
    t = ray.New(40, 30)
    assert (t.width, t.height) == (40, 30) and t.imageData.shape == (30, 40, 4) and not t.imageData.any()


def test_tracer_defaults_after_prepare():
    t = ray.New(20, 10)
    t._prepare(ray.DefaultScene())
    assert (t.FocalLength, t.VerticalFoV, t.MaxDepth, t.NumRaysPerPixel, t.RayRadius) == (1.0, 90.0, 10, 1, 0.5)
    assert t.NumWorkers == (os.cpu_count() or 1) and t.FocusDistance == 1.0 and t.Up == (0.0, 1.0, 0.0)
    t2 = ray.New(20, 10)
    t2.MaxDepth, t2.NumRaysPerPixel, t2.RayRadius, t2.NumWorkers = 7, 3, 0.25, 2
    t2._prepare(ray.DefaultScene())
    assert (t2.MaxDepth, t2.NumRaysPerPixel, t2.RayRadius, t2.NumWorkers) == (7, 3, 0.25, 2)


def test_nil_scene_camera_and_default_background_mutation():
    t = ray.New(8, 8)
    sc = t._prepare(None)  # ray/tracer.go:49-61
    assert len(sc.Objects) == 5 and t.Position == (-2.0, 2.0, 1.0) and t.VerticalFoV == 20.0 and t.Aperture == 0.1
    assert t.FocusDistance == math.sqrt(4 + 4 + 4)
    empty = ray.Scene()
    ray.New(4, 4)._prepare(empty)  # mutates the caller's scene like tracer.go:63-65
    assert empty.Background.ColorA == (1.0, 1.0, 1.0) and empty.Background.ColorB == (0.4, 0.65, 1.0)


# ---- Camera, ray/camera_test.go --------------------------------------------------------------------
def test_camera_position_equals_lookat_does_not_crash():
    c = ray.Camera(Position=(1, 1, 1), LookAt=(1, 1, 1))
    c.Initialize(10, 10)
    assert all(math.isfinite(v) for v in c.pixel00 + c.pixelXVector + c.pixelYVector)


def test_camera_defaults():
    c = ray.Camera()
    c.Initialize(100, 50)
    assert (c.FocalLength, c.VerticalFoV, c.FocusDistance, c.LookAt, c.Up) == (1.0, 90.0, 1.0, (0.0, 0.0, -1.0), (0.0, 1.0, 0.0))
    c2 = ray.Camera(FocalLength=5)
    c2.Initialize(10, 10)
    assert c2.FocusDistance == 5.0  # camera_test.go:164-175


def test_camera_matches_oracle_bit_for_bit(O):
    for (w, h), kw in (((400, 225), O.RICH_CAMERA), ((1920, 1080), O.RICH_CAMERA), ((3840, 2160), O.RICH_CAMERA),
                       ((10, 10), dict(focal_length=5, vfov=30.0)), ((7, 13), dict(position=(-2, 2, 1), look_at=(0, 0, -1), vfov=20.0, aperture=.1, focus_distance=3.4)),
                       # the default 90 degrees (Go's Tan gives exactly 1 at Pi/4, libm 1 - 2^-53) and angles beyond it
                       ((640, 360), dict()), ((33, 17), dict(position=(1, 2, 3), vfov=90.0)), ((64, 36), dict(vfov=100.0, focal_length=2)),
                       ((64, 36), dict(vfov=137.5, look_at=(0.5, -0.25, -1))), ((50, 50), dict(vfov=1.0))):
        oc = O.camera_init(w, h, **kw)
        c = ray.Camera(Position=kw.get("position", (0, 0, 0)), LookAt=kw.get("look_at", (0, 0, 0)), Up=kw.get("up", (0, 0, 0)),
                       VerticalFoV=kw.get("vfov", 0), FocalLength=kw.get("focal_length", 0), FocusDistance=kw.get("focus_distance", 0),
                       Aperture=kw.get("aperture", 0))
        c.Initialize(w, h)
        cc = c.to_c()
        for k in ("position", "pixel00", "pixel_x", "pixel_y", "defocus_u", "defocus_v"):
            assert list(getattr(cc, k)) == list(getattr(oc, k)), (w, h, k)
        assert (cc.aperture, cc.focus_distance, cc.focal_length) == (oc.aperture, oc.focus_distance, oc.focal_length)


def test_go_tan_is_the_oracles_and_not_libm(O):
    """Camera.Initialize calls Go's math.Tan (ray/camera.go:85; pure-Go Cephes form on amd64/arm64): the host mirror restates it
    and agrees with the oracle's restatement bit for bit; libm's tan differs in the last ulp for a quarter of the arguments,
    the default field of view among them."""
    import ctypes
    import math
    import random
    lib = O.lib()
    lib.oracle_go_tan.restype, lib.oracle_go_tan.argtypes = ctypes.c_double, [ctypes.c_double]
    rnd = random.Random(3)
    xs = [rnd.uniform(-3.2, 3.2) for _ in range(20000)] + [rnd.uniform(0, 1.6) for _ in range(20000)] + [0.0, math.pi / 4, -math.pi / 4, 1e-9, 1.5707]
    differ = 0
    for x in xs:
        a, b = ray.go_tan(x), lib.oracle_go_tan(x)
        assert a == b, x
        differ += a != math.tan(x)
    assert ray.go_tan(90.0 * (math.pi / 180.0) / 2.0) == 1.0
    assert differ > 0.1 * len(xs)


def test_cpp_host_go_tan_matches(tmp_path):
    """The C++ host mirror (tray_b200/host/tray.hpp) carries the same restatement of Go's Tan: compiled here and compared with
    the Python one on a spread of arguments (printed as hex floats)."""
    import math
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "tray_b200")
    if not os.path.exists(os.path.join(lib_dir, "libtraycuda.so")):
        pytest.skip("libtraycuda.so not built")
    xs = [0.0, math.pi / 4, -math.pi / 4, 10.0 * (math.pi / 180.0), 0.5, 1.0, 1.2, 1.5707, 2.0, 3.0, -2.5, 1e-9, 0.008726646259971648]
    src = tmp_path / "t.cpp"
    src.write_text('#include "%s/host/tray.hpp"\n#include <cstdio>\nint main() { const double xs[] = {%s};\n'
                   'for (double x : xs) std::printf("%%a\\n", ray::GoTan(x)); return 0; }\n' % (lib_dir, ", ".join(float(x).hex() for x in xs)))
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), str(src), "-L" + lib_dir, "-ltraycuda", "-Wl,-rpath," + lib_dir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [float.fromhex(o) for o in out] == [ray.go_tan(x) for x in xs]


def test_pixel_center_ray_direction():
    # camera_test.go:177-216: ray through pixel (5,5) = pixel00 + 5*px + 5*py - position
    c = ray.Camera(Position=(0, 0, 0), LookAt=(0, 0, -1), VerticalFoV=90.0)
    c.Initialize(10, 10)
    d = ray.Sub(ray.Add(ray.Add(c.pixel00, ray.SMul(c.pixelXVector, 5.0)), ray.SMul(c.pixelYVector, 5.0)), c.Position)
    assert abs(d[0] - 0.1) < 1e-12 and abs(d[1] + 0.1) < 1e-12 and abs(d[2] + 1.0) < 1e-12


# ---- scenes + marshalling --------------------------------------------------------------------------
def test_rich_scene_matches_oracle(O):
    for seed, half in ((2, 11), (7, 11), (42, 11), (2, 20)):
        f = ray.RichScene(rand.New(seed), half).flatten()
        o = O.rich_scene(seed, half)
        for k in ("cx", "cy", "cz", "r", "kind", "params"):
            assert np.array_equal(f[k], getattr(o, k)), (seed, k)


def test_host_rand_matches_go_known_answer():
    r = rand.NewIdx(1, 2)
    assert [r.Uint64() for _ in range(3)] == [0xc4f5a58656eef510, 0x9dcec3ad077dec6c, 0xc8d04605312f8088]
    a, b = rand.New(0), rand.New(0)  # seed 0 randomizes each time (main.go:47)
    assert a.state != b.state


def test_flatten_nested_and_unsupported():
    inner = ray.Scene([ray.Sphere((1, 0, 0), 1, ray.Metal((.5, .5, .5), .1)), ray.Sphere((2, 0, 0), 1, ray.Dielectric(1.5))])
    outer = ray.Scene([ray.Sphere((0, 0, 0), 1, ray.Lambertian((.1, .2, .3))), inner, ray.Sphere((3, 0, 0), 2, ray.Lambertian((1, 1, 1)))])
    f = outer.flatten()  # Scene satisfies Hittable (objects.go:28-46): flattened in order
    assert f["cx"].tolist() == [0, 1, 2, 3] and f["kind"].tolist() == [0, 1, 2, 0] and f["params"][1].tolist() == [.5, .5, .5, .1]
    with pytest.raises(ray.TrayError):
        ray.Scene([object()]).flatten()
    with pytest.raises(ray.TrayError):
        ray.Scene([ray.Sphere((0, 0, 0), 1, "plastic")]).flatten()


def test_shard_rows_partition():
    for h, y0, y1 in ((2160, 0, 2160), (225, 0, 225), (37, 5, 30)):
        for n in (1, 2, 3, 4, 8):
            parts = [ray.shard_rows(y0, y1, i, n) for i in range(n)]
            allrows = sorted(r for p in parts for r in p)
            assert allrows == list(range(y0, y1))
            if n > 1 and y1 - y0 >= 8 * n * 4:
                assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 8

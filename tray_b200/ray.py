"""Host-side mirror of the reference's Go package `ray` for the path-tracing hot path.

Same names, argument meaning, defaults and side effects as the reference (fortio/tray, `ray/*.go`),
so tests read like the reference's own; everything per-pixel runs in libtraycuda.so on a B200.
What stays on the host is exactly what SURVEY.md section 8b keeps in Go: scene construction
(RichScene / DefaultScene), Camera.Initialize, Tracer defaulting, and the marshalling of the scene
into the flat SoA the C ABI takes. There is no CPU rendering path here.
"""
import ctypes as C
import math
import secrets
import threading

import numpy as np

from . import _lib
from . import rand  # noqa: F401  (fortio.org/rand host side: rand.New / rand.NewIdx)
from ._lib import LAYOUT_AUTO, LAYOUT_PLAIN, LAYOUT_REGROUP, LAYOUT_WAVEFRONT, SUMS_ACCUMULATE, SUMS_OFF, SUMS_OVERWRITE  # noqa: F401
from ._lib import ACCEL_AUTO, ACCEL_BRUTE, ACCEL_BVH, ACCEL_CLUSTER, FP32, FP64_FMA, FP64_STRICT, FP64_STRICT_BRUTE, SPLIT_SAMPLES, SPLIT_TILES, STREAM_PER_SAMPLE, STREAM_REFERENCE, TrayError  # noqa: F401

# ------------------------------------------------------------------------------------------------
# Vec3 helpers (ray/vec3.go) on plain 3-tuples; op order as in the reference
# ------------------------------------------------------------------------------------------------
def Add(u, v): return (v[0] + u[0], v[1] + u[1], v[2] + u[2])
def Sub(u, v): return (u[0] - v[0], u[1] - v[1], u[2] - v[2])
def SMul(v, t): return (v[0] * t, v[1] * t, v[2] * t)
def SDiv(v, t): return (v[0] / t, v[1] / t, v[2] / t)
def Mul(u, v): return (u[0] * v[0], u[1] * v[1], u[2] * v[2])
def Dot(u, v): return u[0] * v[0] + u[1] * v[1] + u[2] * v[2]
def Cross(u, v): return (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])
def LengthSquared(v): return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]
def Length(v): return math.sqrt(LengthSquared(v))
def Neg(v): return (-v[0], -v[1], -v[2])


def Unit(v):
    l = Length(v)
    return (v[0] / l, v[1] / l, v[2] / l)


def NearZero(v):
    s = 1e-8
    return abs(v[0]) < s and abs(v[1]) < s and abs(v[2]) < s


# ------------------------------------------------------------------------------------------------
# Materials, objects, scene (ray/materials.go, ray/objects.go)
# ------------------------------------------------------------------------------------------------
class Lambertian:
    def __init__(self, Albedo): self.Albedo = tuple(float(c) for c in Albedo)


class Metal:
    def __init__(self, Albedo, Fuzz=0.0):
        self.Albedo = tuple(float(c) for c in Albedo)
        self.Fuzz = float(Fuzz)


class Dielectric:
    def __init__(self, RefIdx): self.RefIdx = float(RefIdx)


class Sphere:
    def __init__(self, Center, Radius, Mat):
        self.Center = tuple(float(c) for c in Center)
        self.Radius = float(Radius)
        self.Mat = Mat


class AmbientLight:
    def __init__(self, ColorA=(0.0, 0.0, 0.0), ColorB=(0.0, 0.0, 0.0)):
        self.ColorA = tuple(float(c) for c in ColorA)
        self.ColorB = tuple(float(c) for c in ColorB)


def DefaultBackground():  # ray/objects.go:106-110
    return AmbientLight((1.0, 1.0, 1.0), (0.4, 0.65, 1.0))


class Scene:
    def __init__(self, Objects=None, Background=None):
        self.Objects = list(Objects) if Objects else []
        self.Background = Background if Background is not None else AmbientLight()

    def flatten(self):
        """Go-side marshalling: walk Objects in order, nested Scenes flattened in place, into
        the SoA arrays of tray_scene_desc. Anything that is not a Sphere with one of the three materials is
        an error -- there is no CPU fallback to hand it to."""
        cx, cy, cz, r, kind, prm = [], [], [], [], [], []

        def walk(objs):
            for o in objs:
                if isinstance(o, Scene):
                    walk(o.Objects)
                    continue
                if not isinstance(o, Sphere):
                    raise TrayError(_lib.E_UNSUPPORTED, "unsupported Hittable %r (GPU backend handles *Sphere only)" % type(o).__name__)
                m = o.Mat
                if isinstance(m, Lambertian):
                    kind.append(_lib.MAT_LAMBERTIAN); prm.append((*m.Albedo, 0.0))
                elif isinstance(m, Metal):
                    kind.append(_lib.MAT_METAL); prm.append((*m.Albedo, m.Fuzz))
                elif isinstance(m, Dielectric):
                    kind.append(_lib.MAT_DIELECTRIC); prm.append((m.RefIdx, 0.0, 0.0, 0.0))
                else:
                    raise TrayError(_lib.E_UNSUPPORTED, "unsupported Material %r" % type(m).__name__)
                cx.append(o.Center[0]); cy.append(o.Center[1]); cz.append(o.Center[2]); r.append(o.Radius)

        walk(self.Objects)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        return dict(cx=f(cx), cy=f(cy), cz=f(cz), r=f(r), kind=np.ascontiguousarray(kind, dtype=np.uint8),
                    params=f(prm).reshape(-1, 4), bg_a=self.Background.ColorA, bg_b=self.Background.ColorB)


def DefaultScene():  # ray/objects.go:112-130
    ground = Lambertian((0.7, 0.8, 0.1))
    center = Lambertian((0.1, 0.2, 0.5))
    left = Dielectric(1.5)
    bubble = Dielectric(1.0 / 1.5)
    right = Metal((1, .8, .8), 0.05)
    return Scene([Sphere((0, 0, -1.2), 0.5, center), Sphere((0, -100.5, -1), 100, ground),
                  Sphere((-1.0, 0, -1), 0.5, left), Sphere((-1.0, 0, -1), 0.4, bubble),
                  Sphere((1.0, 0, -1), 0.5, right)], DefaultBackground())


def RichScene(rng, half=11):
    """ray/objects.go:132-175. `half` generalises the 22x22 grid (half=50 is BASELINE config 4)."""
    world = Scene()
    world.Objects.append(Sphere((0, -1000, 0), 1000, Lambertian((0.5, 0.5, 0.5))))
    for a in range(-half, half):
        for b in range(-half, half):
            chooseMat = rng.Float64()
            center = (float(a) + 0.9 * rng.Float64(), 0.2, float(b) + 0.9 * rng.Float64())
            if Length(Sub(center, (4.0, 0.2, 0.0))) > 0.9:
                if chooseMat < 0.8:
                    albedo = Mul(rng.Vec3(), rng.Vec3())
                    world.Objects.append(Sphere(center, 0.2, Lambertian(albedo)))
                elif chooseMat < 0.95:
                    albedo = (rng.Float64Range(0.5, 1.0), rng.Float64Range(0.5, 1.0), rng.Float64Range(0.5, 1.0))
                    fuzz = rng.Float64() * 0.5
                    world.Objects.append(Sphere(center, 0.2, Metal(albedo, fuzz)))
                else:
                    world.Objects.append(Sphere(center, 0.2, Dielectric(1.5)))
    world.Objects.append(Sphere((0, 1, 0), 1.0, Dielectric(1.5)))
    world.Objects.append(Sphere((-4, 1, 0), 1.0, Lambertian((0.4, 0.2, 0.1))))
    world.Objects.append(Sphere((4, 1, 0), 1.0, Metal((0.7, 0.6, 0.5), 0.0)))
    return world


# ------------------------------------------------------------------------------------------------
# Camera (ray/camera.go)
# ------------------------------------------------------------------------------------------------
def go_tan(x):
    """Go math.Tan (src/math/tan.go: pure-Go Cephes form, no assembly on amd64 / arm64) for |x| < 2**29 -- what
    Camera.Initialize calls (ray/camera.go:85). Only + - * / on doubles, so it is the same on every host; libm's tan may
    differ in the last ulp, which would move pixel00 and every camera ray with it."""
    PI4A, PI4B, PI4C = 7.85398125648498535156e-1, 3.77489470793079817668e-8, 2.69515142907905952645e-15
    P0, P1, P2 = -1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7
    Q1, Q2, Q3, Q4 = 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7, -5.38695755929454629881e7
    if x == 0 or x != x:
        return x
    if math.isinf(x):
        return math.nan
    sign = x < 0
    if sign:
        x = -x
    j = int(x * float.fromhex("0x1.45f306dc9c883p+0"))  # 4/Pi, folded like Go folds the constant
    y = float(j)
    if j & 1:
        j += 1
        y += 1.0
    z = ((x - y * PI4A) - y * PI4B) - y * PI4C
    zz = z * z
    if zz > 1e-14:
        y = z + z * (zz * (((P0 * zz) + P1) * zz + P2) / ((((zz + Q1) * zz + Q2) * zz + Q3) * zz + Q4))
    else:
        y = z
    if j & 2:
        y = -1 / y
    return -y if sign else y


class Camera:
    def __init__(self, Position=(0.0, 0.0, 0.0), LookAt=(0.0, 0.0, 0.0), Up=(0.0, 0.0, 0.0), VerticalFoV=0.0,
                 FocalLength=0.0, FocusDistance=0.0, Aperture=0.0):
        self.Position = tuple(float(c) for c in Position)
        self.LookAt = tuple(float(c) for c in LookAt)
        self.Up = tuple(float(c) for c in Up)
        self.VerticalFoV, self.FocalLength = float(VerticalFoV), float(FocalLength)
        self.FocusDistance, self.Aperture = float(FocusDistance), float(Aperture)
        self.pixel00 = self.pixelXVector = self.pixelYVector = (0.0, 0.0, 0.0)
        self.defocusDiskU = self.defocusDiskV = (0.0, 0.0, 0.0)

    def copy_camera_from(self, c):
        for k in ("Position", "LookAt", "Up", "VerticalFoV", "FocalLength", "FocusDistance", "Aperture"):
            setattr(self, k, getattr(c, k))

    def Initialize(self, width, height):
        """ray/camera.go:43-105, same defaults and evaluation order (float64, no fused ops)."""
        zero = (0.0, 0.0, 0.0)
        if self.FocalLength == 0: self.FocalLength = 1.0
        if self.VerticalFoV == 0: self.VerticalFoV = 90.0
        if tuple(self.Up) == zero: self.Up = (0.0, 1.0, 0.0)
        if self.FocusDistance == 0: self.FocusDistance = self.FocalLength
        if tuple(self.Position) == zero and tuple(self.LookAt) == zero: self.LookAt = (0.0, 0.0, -1.0)
        view = Sub(self.Position, self.LookAt)
        if NearZero(view): view = (0.0, 0.0, 1.0)
        w = Unit(view)
        u = Unit(Cross(self.Up, w))
        v = Cross(w, u)
        defocusRadius = self.Aperture / 2
        self.defocusDiskU = SMul(u, defocusRadius)
        self.defocusDiskV = SMul(v, defocusRadius)
        theta = self.VerticalFoV * (math.pi / 180.0)
        viewportHeight = 2.0 * self.FocalLength * go_tan(theta / 2.0)  # math.Tan as Go computes it (Cephes form), not libm's
        aspectRatio = float(width) / float(height)
        viewportWidth = aspectRatio * viewportHeight
        horizontal = SMul(u, viewportWidth)
        vertical = SMul(v, -viewportHeight)
        self.pixelXVector = SDiv(horizontal, float(width))
        self.pixelYVector = SDiv(vertical, float(height))
        # Position.Minus(a, b, c) = Position - ((a + b) + c)   (ray/vec3.go:44-55)
        upperLeft = Sub(self.Position, Add(Add(SMul(w, self.FocalLength), SMul(horizontal, 0.5)), SMul(vertical, 0.5)))
        self.pixel00 = Add(upperLeft, SMul(Add(self.pixelXVector, self.pixelYVector), 0.5))

    def to_c(self):
        c = _lib.CameraC()
        c.position[:] = self.Position
        c.pixel00[:] = self.pixel00
        c.pixel_x[:] = self.pixelXVector
        c.pixel_y[:] = self.pixelYVector
        c.defocus_u[:] = self.defocusDiskU
        c.defocus_v[:] = self.defocusDiskV
        c.aperture, c.focus_distance, c.focal_length = self.Aperture, self.FocusDistance, self.FocalLength
        return c


def RichSceneCamera():  # ray/camera.go:144-154
    return Camera(Position=(13, 2, 3), LookAt=(0, 0, 0), Up=(0, 1, 0), VerticalFoV=20.0, Aperture=0.1,
                  FocalLength=10.0, FocusDistance=10.0)


# ------------------------------------------------------------------------------------------------
# Device context (one per process, shared by tracers unless told otherwise)
# ------------------------------------------------------------------------------------------------
class Context:
    """Owns a tray_ctx (device memory, streams). `devices`: list of CUDA ordinals (default [0])."""

    def __init__(self, devices=None):
        L = _lib.lib()
        self._L = L
        self.handle = C.c_void_p()
        n = len(devices) if devices else 0
        arr = (C.c_int * n)(*devices) if n else None
        rc = L.tray_init(arr, n, C.byref(self.handle))
        if rc != 0:
            raise TrayError(rc, (L.tray_last_error(None) or b"?").decode())
        self.devices = list(devices) if devices else [0]
        self._scene_key = None

    def close(self):
        if self.handle:
            self._L.tray_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, flat):
        d = _lib.SceneDesc()
        d.n = len(flat["cx"])
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        d.cx, d.cy, d.cz, d.radius = p(flat["cx"]), p(flat["cy"]), p(flat["cz"]), p(flat["r"])
        d.mat_kind, d.mat_params = p(flat["kind"]), p(flat["params"])
        d.bg_a[:] = flat["bg_a"]
        d.bg_b[:] = flat["bg_b"]
        _lib.check(self.handle, self._L.tray_scene_upload(self.handle, C.byref(d)))

    def render(self, cam_c, params, out=None):
        st = _lib.Stats()
        ptr, stride = (None, 0)
        if out is not None:
            ptr, stride = out.ctypes.data_as(C.c_void_p), out.strides[0]
        _lib.check(self.handle, self._L.tray_render(self.handle, C.byref(cam_c), C.byref(params), ptr, stride, C.byref(st)))
        return st.as_dict()

    def read_image(self, out):
        _lib.check(self.handle, self._L.tray_read_image(self.handle, out.ctypes.data_as(C.c_void_p), out.strides[0]))

    def read_hdr(self, width, height):
        hdr = np.zeros((height, width, 3))
        _lib.check(self.handle, self._L.tray_read_hdr(self.handle, hdr.ctypes.data_as(C.c_void_p)))
        return hdr

    def device_sums(self):
        """(device pointer, number of doubles) of the raw colour sums left by the last SUMS_* render."""
        ptr, n = C.c_void_p(), C.c_uint64()
        _lib.check(self.handle, self._L.tray_device_sums(self.handle, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def resolve_sums(self, n_samples=0, out=None):
        ptr, stride = (None, 0)
        if out is not None:
            ptr, stride = out.ctypes.data_as(C.c_void_p), out.strides[0]
        _lib.check(self.handle, self._L.tray_resolve_sums(self.handle, n_samples, ptr, stride))

    def first_hit(self, cam_c, width, height, precision=FP64_STRICT):
        ids = np.zeros((height, width), dtype=np.int32)
        t = np.zeros((height, width))
        nrm = np.zeros((height, width, 3))
        front = np.zeros((height, width), dtype=np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(self.handle, self._L.tray_first_hit(self.handle, C.byref(cam_c), width, height, precision,
                                                       p(ids), p(t), p(nrm), p(front)))
        return ids, t, nrm, front

    def rng_dump(self, kind, idx, seed, n, radius=1.0):
        per = {0: 1, 1: 1, 2: 1, 3: 3, 4: 2}[kind]
        out = np.zeros(n * per)
        _lib.check(self.handle, self._L.tray_rng_dump(self.handle, kind, idx, seed, radius, n, out.ctypes.data_as(C.c_void_p)))
        if kind == 0:
            return out.view(np.uint64)
        return out.reshape(n, per) if per > 1 else out

    def arith_probe(self, kind, a, b):
        """Device IEEE division / sqrt routines on given operands (kind: 0 div3, 1 shared-reciprocal division, 2 sqrt, 3 plain division)."""
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        out = np.zeros(len(a) * (3 if kind == 0 else 1))
        p = lambda v: v.ctypes.data_as(C.c_void_p)
        _lib.check(self.handle, self._L.tray_arith_probe(self.handle, kind, p(a), p(b), len(a), p(out)))
        return out.reshape(len(a), 3) if kind == 0 else out

    def linear_to_srgb(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(len(x), dtype=np.uint8)
        _lib.check(self.handle, self._L.tray_linear_to_srgb(self.handle, x.ctypes.data_as(C.c_void_p), len(x), out.ctypes.data_as(C.c_void_p)))
        return out

    def present(self, cols, rows2, want_image=True):
        """main.go:119-130 on the device: BiLinear downscale of the last frame to cols x rows2 and the half-block ANSI frame
        for a cols x rows2/2 terminal. Returns (ansi bytes, small image or None, device ms)."""
        n = C.c_size_t()
        ms = C.c_double()
        small = np.zeros((rows2, cols, 4), dtype=np.uint8) if want_image else None
        cap = (rows2 // 2) * (cols * 41 + 5)
        buf = np.zeros(cap, dtype=np.uint8)
        _lib.check(self.handle, self._L.tray_present(self.handle, cols, rows2, small.ctypes.data_as(C.c_void_p) if want_image else None,
                                                     buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(ms)))
        return buf[:n.value].tobytes(), small, ms.value

    def upload_frame(self, img):
        """An arbitrary (h, w, 4) uint8 image becomes "the last rendered frame" (for present / encode_png / tests)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape[:2]
        _lib.check(self.handle, self._L.tray_upload_frame(self.handle, img.ctypes.data_as(C.c_void_p), img.strides[0], w, h))

    def encode_png(self, width, height):
        """SaveImage's png.Encode (main.go:26-36) on the device: returns (PNG file bytes, device ms)."""
        cap = int(self._L.tray_png_bound(width, height))
        buf = np.empty(cap, dtype=np.uint8)
        n, ms = C.c_size_t(), C.c_double()
        _lib.check(self.handle, self._L.tray_encode_png(self.handle, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.byref(ms)))
        return buf[:n.value].tobytes(), ms.value

    def configure(self, key, value):
        _lib.check(self.handle, self._L.tray_configure(self.handle, key, value))

    def query(self, key):
        return int(self._L.tray_query(self.handle, key))

    def progress(self):
        return int(self._L.tray_progress(self.handle))

    def measure_peak(self, kind=0):
        tf, ms = C.c_double(), C.c_double()
        _lib.check(self.handle, self._L.tray_measure_peak(self.handle, kind, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


# ------------------------------------------------------------------------------------------------
# Tracer (ray/tracer.go)
# ------------------------------------------------------------------------------------------------
class Tracer(Camera):
    """ray.Tracer: embeds Camera; exported fields MaxDepth, NumRaysPerPixel, RayRadius, NumWorkers,
    ProgressFunc, Seed (ray/tracer.go:25-35). Additive backend knobs: StreamMode, Precision, SplitMode,
    ShardIndex/ShardCount, Context."""

    def __init__(self, width, height):
        Camera.__init__(self)
        self.MaxDepth = 0
        self.NumRaysPerPixel = 0
        self.RayRadius = 0.0
        self.NumWorkers = 0
        self.ProgressFunc = None
        self.Seed = 0
        self.width, self.height = int(width), int(height)
        self.imageData = np.zeros((self.height, self.width, 4), dtype=np.uint8)  # image.NewRGBA: zeroed
        # backend
        self.StreamMode = STREAM_PER_SAMPLE
        self.Precision = FP64_STRICT  # exact Go/amd64 float64 semantics; FP64_FMA is the opt-in fused mode
        self.SplitMode = SPLIT_TILES
        self.Accel = ACCEL_AUTO
        self.Layout = LAYOUT_AUTO
        self.ShardIndex, self.ShardCount = 0, 0
        self.Context = None
        self.Stats = None
        self._seed_used = 0

    @property
    def Camera(self):
        return self

    @Camera.setter
    def Camera(self, cam):
        self.copy_camera_from(cam)

    def _prepare(self, scene):
        # nil scene -> DefaultScene + hard-coded camera (ray/tracer.go:49-61)
        if scene is None:
            scene = DefaultScene()
            self.Position = (-2.0, 2.0, 1.0)
            self.LookAt = (0.0, 0.0, -1.0)
            self.VerticalFoV = 20.0
            self.Aperture = .1
            self.FocusDistance = Length(Sub(self.Position, self.LookAt))
        # default background; mutates the caller's scene like the reference (tracer.go:63-65)
        zero = (0.0, 0.0, 0.0)
        if tuple(scene.Background.ColorA) == zero and tuple(scene.Background.ColorB) == zero:
            scene.Background = DefaultBackground()
        if self.MaxDepth <= 0: self.MaxDepth = 10
        if self.NumRaysPerPixel <= 0: self.NumRaysPerPixel = 1
        if self.RayRadius <= 0: self.RayRadius = 0.5
        if self.NumWorkers <= 0:
            import os
            self.NumWorkers = os.cpu_count() or 1  # runtime.GOMAXPROCS(0)
        self.Initialize(self.width, self.height)
        return scene

    def _params(self, y0, y1, stream_idx=-1):
        p = _lib.Params()
        p.width, p.height = self.width, self.height
        p.spp, p.max_depth, p.ray_radius = self.NumRaysPerPixel, self.MaxDepth, self.RayRadius
        seed = self.Seed
        if seed == 0:  # randomized each time (ray/tracer.go:32)
            seed = secrets.randbits(64) | 1
        self._seed_used = seed
        p.seed = seed
        p.y0, p.y1 = y0, y1
        p.stream_mode, p.num_workers, p.stream_idx = self.StreamMode, self.NumWorkers, stream_idx
        p.precision, p.split_mode, p.accel, p.layout = self.Precision, self.SplitMode, self.Accel, self.Layout
        p.shard_index, p.shard_count = self.ShardIndex, self.ShardCount
        return p

    def _run(self, scene, params):
        ctx = self.Context or default_context()
        ctx.upload(scene.flatten())
        cam_c = self.to_c()
        total = (params.y1 - params.y0) * self.width
        if self.ProgressFunc is None:
            self.Stats = ctx.render(cam_c, params, self.imageData)
            return
        # ProgressFunc: poll the library from a helper thread; deltas sum to the pixel count
        # (ray/tracer_test.go:172-186). The reference also calls it from worker goroutines.
        done = threading.Event()
        sent = [0]

        def poll():
            while not done.wait(0.02):
                cur = min(ctx.progress(), total)
                if cur > sent[0]:
                    self.ProgressFunc(cur - sent[0])
                    sent[0] = cur

        th = threading.Thread(target=poll, daemon=True)
        th.start()
        try:
            self.Stats = ctx.render(cam_c, params, self.imageData)
        finally:
            done.set()
            th.join()
        if total > sent[0]:
            self.ProgressFunc(total - sent[0])

    def Render(self, scene):
        """(*Tracer).Render (ray/tracer.go:48-118): returns the tracer's own imageData."""
        scene = self._prepare(scene)
        self._run(scene, self._params(0, self.height))
        return self.imageData

    def RenderProgressive(self, scene, slice_rays):
        """Additive (not in the reference, whose README lists navigation as WIP): the same frame as Render, delivered as
        a sequence of refinements. Yields (rays_done, imageData) after every `slice_rays` rays per pixel; the scene and
        the partial sums stay on the device, each slice continues the per-pixel sum in sample order, so the last image
        is bit-identical to Render's."""
        scene = self._prepare(scene)
        ctx = self.Context or default_context()
        ctx.upload(scene.flatten())
        cam_c = self.to_c()
        n, done = self.NumRaysPerPixel, 0
        seed = self.Seed or (secrets.randbits(64) | 1)
        stats = None
        while done < n:
            k = min(int(slice_rays), n - done)
            p = self._params(0, self.height)
            p.seed = seed
            p.sample_offset, p.sample_stride, p.sample_count = done, 1, k
            p.sums_mode = SUMS_OVERWRITE if done == 0 else SUMS_ACCUMULATE
            st = ctx.render(cam_c, p, None)
            if stats is None:
                stats = dict(st)
            else:
                for key in ("paths", "segments", "sphere_tests", "depth_exhausted", "kernel_ms", "total_ms", "launches", "trace_kernel_ms"):
                    stats[key] += st[key]
            done += k
            ctx.resolve_sums(done, self.imageData)
            self.Stats = stats
            yield done, self.imageData

    def RenderLines(self, idx, yStart, yEnd, scene):
        """(*Tracer).RenderLines (ray/tracer.go:120-155): rows [yStart,yEnd) only; idx = stream index
        (used in STREAM_REFERENCE mode; per-sample streams do not depend on it)."""
        if self.NumRaysPerPixel <= 0 or self.MaxDepth <= 0:
            raise TrayError(_lib.E_INVALID, "RenderLines needs defaulted parameters (call Render, or set them)")
        p = self._params(yStart, yEnd, stream_idx=idx)
        p.num_workers = 0
        self._run(scene, p)


def SaveImage(tracer, fname):
    """main.go:26-36 / benchmark/benchmark.go:23-33 (SaveImage(img, fname) -> png.Encode): writes the tracer's last frame
    as PNG. The file is encoded on the device from the frame still resident in HBM; returns the encode time in ms."""
    ctx = tracer.Context or default_context()
    data, ms = ctx.encode_png(tracer.width, tracer.height)
    with open(fname, "wb") as f:
        f.write(data)
    return ms


def New(width, height):
    """ray.New (ray/tracer.go:38-45)."""
    return Tracer(width, height)


BAND_ROWS = 1  # kBandRows in csrc/tray_api.cu


def shard_rows(y0, y1, shard_index, shard_count, band_rows=BAND_ROWS):
    """Rows of [y0,y1) that shard `shard_index` of `shard_count` renders in tile mode: rows, round-robin
    (same rule as tray_render; the reference's analogue is the row-chunk work queue, ray/tracer.go:93-103)."""
    if shard_count <= 1:
        return list(range(y0, y1))
    return [y0 + r for r in range(y1 - y0) if (r // band_rows) % shard_count == shard_index]

"""ctypes binding of libtraycuda.so -- exactly the entry points include/tray_cuda.h declares."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("TRAY_LIB") or os.path.join(_HERE, "libtraycuda.so")  # TRAY_LIB: tuning variants only

# error codes / enums (include/tray_cuda.h)
OK, E_INVALID, E_CUDA, E_NO_SCENE, E_UNSUPPORTED, E_NO_DEVICE = 0, -1, -2, -3, -4, -5
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC = 0, 1, 2
STREAM_REFERENCE, STREAM_PER_SAMPLE = 0, 1
FP64_FMA, FP64_STRICT, FP32, FP64_STRICT_BRUTE = 0, 1, 2, 3
SPLIT_TILES, SPLIT_SAMPLES = 0, 1
ACCEL_AUTO, ACCEL_BRUTE, ACCEL_BVH, ACCEL_CLUSTER = 0, 1, 2, 3
SUMS_OFF, SUMS_OVERWRITE, SUMS_ACCUMULATE = 0, 1, 2
LAYOUT_AUTO, LAYOUT_PLAIN, LAYOUT_REGROUP, LAYOUT_WAVEFRONT = 0, 1, 2, 3
CFG_BVH_BUILD, BVH_BUILD_AUTO, BVH_BUILD_HOST, BVH_BUILD_DEVICE = 1, 0, 1, 2
CFG_INDISC, CFG_UNITVEC = 2, 3

EXPORTS = ("tray_init", "tray_destroy", "tray_last_error", "tray_abi_version", "tray_scene_upload", "tray_render",
           "tray_read_image", "tray_read_hdr", "tray_first_hit", "tray_rng_dump", "tray_linear_to_srgb",
           "tray_progress", "tray_measure_peak", "tray_present", "tray_device_sums", "tray_resolve_sums", "tray_png_bound", "tray_encode_png", "tray_configure", "tray_query", "tray_upload_frame",
           "tray_cluster_tables", "tray_arith_probe")


class TrayError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libtraycuda error %d: %s" % (code, msg))
        self.code = code


class SceneDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("cx", C.c_void_p), ("cy", C.c_void_p), ("cz", C.c_void_p), ("radius", C.c_void_p),
                ("mat_kind", C.c_void_p), ("mat_params", C.c_void_p), ("bg_a", C.c_double * 3), ("bg_b", C.c_double * 3)]


class CameraC(C.Structure):
    _fields_ = [(k, C.c_double * 3) for k in ("position", "pixel00", "pixel_x", "pixel_y", "defocus_u", "defocus_v")] + \
               [("aperture", C.c_double), ("focus_distance", C.c_double), ("focal_length", C.c_double)]


class Params(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("ray_radius", C.c_double), ("seed", C.c_uint64), ("y0", C.c_int32), ("y1", C.c_int32),
                ("stream_mode", C.c_int32), ("num_workers", C.c_int32), ("stream_idx", C.c_int64),
                ("precision", C.c_int32), ("split_mode", C.c_int32), ("shard_index", C.c_int32),
                ("shard_count", C.c_int32), ("accel", C.c_int32), ("sample_offset", C.c_int32), ("sample_stride", C.c_int32),
                ("sample_count", C.c_int32), ("sums_mode", C.c_int32), ("layout", C.c_int32), ("reserved", C.c_int32 * 3)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("depth_exhausted", C.c_uint64), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("launches", C.c_int32), ("n_devices", C.c_int32), ("trace_kernel_ms", C.c_double),
                ("trace_launches", C.c_double), ("box_tests", C.c_double), ("bounds_violations", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


def library_path():
    return _SO


def build_library(force=False, verbose=False):
    """Compile libtraycuda.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in ("tray_api.cu", "tray_kernels.cuh", "tray_device.cuh", "tray_png.cuh", "tray_wavefront.cuh", "tray_lbvh.cuh", "zig_tables.h")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "tray_cuda.h"))
    stale = not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    # the C++ host CLIs (the reference's two mains on this backend) are built by the same Makefile
    host = [os.path.join(_HERE, "host", f) for f in ("tray.hpp", "benchmark.cpp", "tray.cpp")]
    for exe in ("benchmark", "tray"):
        e = os.path.join(_HERE, exe)
        stale = stale or not os.path.exists(e) or any(os.path.getmtime(s) > os.path.getmtime(e) for s in host)
    if force or stale:
        cmd = ["make", "-C", os.path.join(_HERE, "csrc")] + (["-B"] if force else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
        if out.returncode != 0:
            raise RuntimeError("building libtraycuda.so failed")
    return _SO


_lib = None


def lib():
    """Loads the CUDA library; fails loudly if it is missing (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise TrayError(E_NO_DEVICE, "libtraycuda.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                         "tray_b200 has no CPU fallback")
        L = C.CDLL(_SO)
        L.tray_init.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
        L.tray_destroy.argtypes = [C.c_void_p]
        L.tray_destroy.restype = None
        L.tray_last_error.argtypes = [C.c_void_p]
        L.tray_last_error.restype = C.c_char_p
        L.tray_abi_version.restype = C.c_int
        L.tray_scene_upload.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        L.tray_render.argtypes = [C.c_void_p, C.POINTER(CameraC), C.POINTER(Params), C.c_void_p, C.c_size_t, C.POINTER(Stats)]
        L.tray_read_image.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.tray_read_hdr.argtypes = [C.c_void_p, C.c_void_p]
        L.tray_first_hit.argtypes = [C.c_void_p, C.POINTER(CameraC), C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 4
        L.tray_rng_dump.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_uint64, C.c_double, C.c_int32, C.c_void_p]
        L.tray_linear_to_srgb.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.tray_arith_probe.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.tray_present.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                   C.POINTER(C.c_double)]
        L.tray_device_sums.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.tray_resolve_sums.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]
        L.tray_png_bound.argtypes = [C.c_int32, C.c_int32]
        L.tray_png_bound.restype = C.c_size_t
        L.tray_encode_png.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_double)]
        L.tray_configure.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
        L.tray_query.argtypes = [C.c_void_p, C.c_int32]
        L.tray_query.restype = C.c_int64
        L.tray_upload_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32]
        L.tray_cluster_tables.argtypes = [C.POINTER(SceneDesc), C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.tray_progress.argtypes = [C.c_void_p]
        L.tray_progress.restype = C.c_uint64
        L.tray_measure_peak.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def check(ctx, rc):
    if rc != 0:
        msg = lib().tray_last_error(ctx)
        raise TrayError(rc, msg.decode() if msg else "?")

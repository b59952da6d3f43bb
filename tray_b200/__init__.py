"""tray_b200 -- B200-native path-tracing backend for fortio/tray.

`tray_b200.ray` mirrors the reference's Go package `ray` (same names and argument meaning) on top of
the C ABI of libtraycuda.so (include/tray_cuda.h). There is no CPU fallback: importing works anywhere,
rendering needs a B200.
"""
from . import ray  # noqa: F401
from ._lib import TrayError, build_library, library_path  # noqa: F401

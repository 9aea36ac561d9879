"""ctypes binding of the C-ABI CUDA library (include/movis_b200.h).

The library is built in-tree (``make -C movis_b200/csrc`` or ``__graft_entry__.build()``).  There
is no fallback: if the shared object is missing, or no sm_100 device is usable, every product call
raises -- this package never computes pixels on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MOVIS_B200_LIB: an experiment build of the same library (movis_b200/csrc/Makefile: `make variant`)
LIB_PATH = os.environ.get("MOVIS_B200_LIB") or os.path.join(_HERE, "csrc", "libmovis_b200.so")

MVB_LAYER_PIL_OVER = 1
MVB_LAYER_OPAQUE_SOURCE = 2
MVB_STACK_KEEP_FRAME = 1
MVB_BLUR_PLAIN, MVB_BLUR_GLOW = 0, 1

# numpy mirror of `struct mvb_layer` (112 bytes, natural alignment)
LAYER_DTYPE = np.dtype([
    ("src", np.uint64), ("src_w", np.int32), ("src_h", np.int32), ("src_stride", np.int64),
    ("M", np.float64, (6,)),
    ("bbox_x", np.int32), ("bbox_y", np.int32), ("bbox_w", np.int32), ("bbox_h", np.int32),
    ("blend_mode", np.int32), ("flags", np.int32), ("opacity", np.float64), ("alpha_map", np.uint64),
], align=True)
assert LAYER_DTYPE.itemsize == 112

# every symbol include/movis_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "mvb_last_error", "mvb_abi_version", "mvb_device_check", "mvb_kernel_launches", "mvb_max_layers_per_launch",
    "mvb_composite_stack", "mvb_composite_frames", "mvb_warp_affine_rgba8", "mvb_alpha_composite",
    "mvb_blur_ksize", "mvb_gaussian_weights_q8", "mvb_effect_blur_scratch_bytes",
    "mvb_effect_blur_rgba8", "mvb_effect_drop_shadow_rgba8", "mvb_effect_chroma_key_rgba8",
    "mvb_effect_fill_color_rgba8", "mvb_effect_hsl_shift_rgba8", "mvb_rgb_to_hsv_u8",
    "mvb_source_stripe_rgba8", "mvb_alpha_map_rgba8",
]


class NativeLibraryError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library; raises ``NativeLibraryError`` when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build the sm_100a CUDA extension first "
            "(`make -C movis_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "movis_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl, u8p = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_void_p
    L.mvb_last_error.restype = C.c_char_p
    L.mvb_abi_version.restype = i32
    L.mvb_device_check.restype = i32
    L.mvb_max_layers_per_launch.restype = i32
    L.mvb_kernel_launches.restype = C.c_uint64
    L.mvb_composite_stack.argtypes = [vp, i32, i32, i64, u8p, i32, vp, i32, vp]
    L.mvb_composite_frames.argtypes = [i32, vp, i32, i32, i64, u8p, vp, vp, i32, vp]
    L.mvb_warp_affine_rgba8.argtypes = [vp, i32, i32, i64, vp, i32, i32, i64, vp, vp]
    L.mvb_alpha_composite.argtypes = [vp, i32, i32, i64, vp, i32, i32, i64, i32, i32, dbl, i32, i32, i32, vp]
    L.mvb_alpha_map_rgba8.argtypes = [vp, i32, i32, i64, vp, vp]
    L.mvb_blur_ksize.argtypes = [dbl]
    L.mvb_blur_ksize.restype = i32
    L.mvb_gaussian_weights_q8.argtypes = [dbl, i32, vp]
    L.mvb_effect_blur_scratch_bytes.argtypes = [i32, i32, i32, i32]
    L.mvb_effect_blur_scratch_bytes.restype = C.c_size_t
    L.mvb_effect_blur_rgba8.argtypes = [vp, i32, i32, i64, vp, i64, i32, vp, i32, vp, vp, vp]
    L.mvb_effect_drop_shadow_rgba8.argtypes = [vp, i32, i32, i64, vp, i64, i32, vp, vp, vp, i32, i32, vp, vp]
    L.mvb_effect_chroma_key_rgba8.argtypes = [vp, i32, i32, i64, vp, i64, vp, vp, vp]
    L.mvb_effect_fill_color_rgba8.argtypes = [vp, i32, i32, i64, vp, i64, vp, vp]
    L.mvb_effect_hsl_shift_rgba8.argtypes = [vp, i32, i32, i64, vp, i64, vp, vp, vp, vp]
    L.mvb_rgb_to_hsv_u8.argtypes = [vp, vp]
    L.mvb_source_stripe_rgba8.argtypes = [vp, i32, i32, i64, vp, i32, vp]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:  # default restype: our int-returning calls
            fn.restype = i32
    _lib = L
    return L


def check(rc: int) -> None:
    """Raise on a negative C-ABI return code."""
    if rc != 0:
        msg = lib().mvb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"movis_b200: {msg}")
        raise RuntimeError(f"movis_b200 (error {rc}): {msg}")


def blur_ksize(radius: float) -> int:
    return int(lib().mvb_blur_ksize(float(radius)))


def gaussian_weights_q8(sigma: float, ksize: int) -> np.ndarray:
    w = np.empty(ksize, dtype=np.uint16)
    check(lib().mvb_gaussian_weights_q8(float(sigma), int(ksize), w.ctypes.data))
    return w


def rgb_to_hsv_u8(rgb) -> np.ndarray:
    src = np.ascontiguousarray(np.asarray(rgb, dtype=np.uint8).reshape(3))
    out = np.empty(3, dtype=np.uint8)
    check(lib().mvb_rgb_to_hsv_u8(src.ctypes.data, out.ctypes.data))
    return out

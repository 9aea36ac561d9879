"""ctypes front-end of the CPU oracle (oracle/mvo.c) plus the numpy restatement of the reference's
host-side geometry.

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; nothing under
``movis_b200/`` does.  Parity of this oracle with the reference is pinned by the golden vectors in
``tests/golden/`` (generated from the live reference by ``oracle/gen_golden.py``).

All ``path:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmvo.so")
_lib = None

BLEND_MODES = [
    "normal", "multiply", "screen", "overlay", "darken", "lighten", "color_dodge", "color_burn",
    "linear_dodge", "linear_burn", "hard_light", "soft_light", "vivid_light", "linear_light",
    "pin_light", "difference", "exclusion", "subtract",
]  # enum.py:173-192, value == index
MATTE_MODES = ["none", "alpha", "luminance"]  # enum.py:225-229


def build(force: bool = False) -> str:
    """Compile oracle/mvo.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "mvo.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", _HERE, "CC=gcc"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


class _Layer(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_w", C.c_int32), ("src_h", C.c_int32), ("src_stride", C.c_int64),
        ("M", C.c_double * 6),
        ("bbox_x", C.c_int32), ("bbox_y", C.c_int32), ("bbox_w", C.c_int32), ("bbox_h", C.c_int32),
        ("blend_mode", C.c_int32), ("use_pil", C.c_int32), ("opacity", C.c_double),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p8, i, d, sz, pd = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_ssize_t
        L.mvo_num_threads.restype = C.c_int
        L.mvo_set_num_threads.argtypes = [i]
        L.mvo_invert_affine.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mvo_warp_affine_rgba8.argtypes = [p8, i, i, pd, p8, i, i, pd, C.POINTER(C.c_double)]
        L.mvo_blend_lut.argtypes = [i, p8]
        L.mvo_alpha_composite.argtypes = [p8, i, i, pd, p8, i, i, pd, i, i, d, i, i, i]
        L.mvo_composite_stack.argtypes = [p8, i, i, pd, p8, i, C.POINTER(_Layer)]
        L.mvo_composite_stack.restype = i
        L.mvo_blur_ksize.argtypes = [d]
        L.mvo_blur_ksize.restype = i
        L.mvo_gaussian_weights_q8.argtypes = [d, i, p8]
        L.mvo_gaussian_blur_u8.argtypes = [p8, i, i, pd, i, i, p8, p8, pd]
        L.mvo_pad_edge_rgb_zero_alpha.argtypes = [p8, i, i, pd, i, p8, pd]
        L.mvo_effect_gaussian_blur.argtypes = [p8, i, i, pd, d, p8, pd]
        L.mvo_effect_glow.argtypes = [p8, i, i, pd, d, p8, p8, pd]
        L.mvo_effect_drop_shadow.argtypes = [p8, i, i, pd, d, p8, p8, i, i, p8, pd]
        L.mvo_hsv_tables.argtypes = [p8, p8]
        L.mvo_rgb2hsv.argtypes = [p8, sz, i, p8]
        L.mvo_hsv2rgb.argtypes = [p8, i, i, p8]
        L.mvo_effect_chroma_key.argtypes = [p8, i, i, pd, p8, p8, p8, pd]
        L.mvo_effect_fill_color.argtypes = [p8, i, i, pd, p8, p8, pd]
        L.mvo_effect_hsl_shift.argtypes = [p8, i, i, pd, p8, p8, p8, p8, pd]
        L.mvo_stripe.argtypes = [p8, i, i, pd, p8, d, d, d, d, p8, p8, i]
        _lib = L
    return _lib


def _rgba(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    assert a.dtype == np.uint8 and a.ndim == 3 and a.shape[2] == 4, (a.dtype, a.shape)
    if a.strides[2] != 1 or a.strides[1] != 4:
        a = np.ascontiguousarray(a)
    return a


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _mode_index(mode) -> int:
    if isinstance(mode, str):
        return BLEND_MODES.index(mode)
    return int(getattr(mode, "value", mode))


def _matte_index(m) -> int:
    if isinstance(m, str):
        return MATTE_MODES.index(m)
    return int(getattr(m, "value", m))


def num_threads() -> int:
    return lib().mvo_num_threads()


def set_num_threads(n: int) -> None:
    lib().mvo_set_num_threads(int(n))


# ---------------------------------------------------------------------------------------------
# pixel operators
# ---------------------------------------------------------------------------------------------

def invert_affine(M: np.ndarray) -> np.ndarray:
    Mf = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(6))
    out = np.empty(6, dtype=np.float64)
    lib().mvo_invert_affine(Mf.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out.reshape(2, 3)


def warp_affine(src: np.ndarray, M: np.ndarray, dsize: tuple[int, int]) -> np.ndarray:
    """cv2.warpAffine(src, M, dsize, INTER_LINEAR, BORDER_CONSTANT) -- composition.py:811-813."""
    src = _rgba(src)
    W, H = int(dsize[0]), int(dsize[1])
    dst = np.empty((H, W, 4), dtype=np.uint8)
    Mf = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(6))
    lib().mvo_warp_affine_rgba8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), W, H,
                                dst.strides[0], Mf.ctypes.data_as(C.POINTER(C.c_double)))
    return dst


def blend_lut(mode) -> np.ndarray:
    """T[b, f] for one blend mode (imgproc.py:11-133)."""
    out = np.empty((256, 256), dtype=np.uint8)
    lib().mvo_blend_lut(_mode_index(mode), _ptr(out))
    return out


def alpha_composite(bg: np.ndarray, fg: np.ndarray, position=(0, 0), opacity: float = 1.0,
                    blending_mode="normal", matte_mode="none", use_pil: bool | None = None) -> np.ndarray:
    """movis.imgproc.alpha_composite (imgproc.py:216-273).  Returns a new array.

    ``use_pil`` selects the PIL fast path; the default mirrors the reference's dispatch for an
    *enum* NORMAL + MatteMode.NONE (a ``str`` 'normal' goes down the numpy path there,
    imgproc.py:265-273 -- pass use_pil=False to model that)."""
    bg = _rgba(bg).copy()
    fg = _rgba(fg)
    mode, matte = _mode_index(blending_mode), _matte_index(matte_mode)
    if use_pil is None:
        use_pil = (mode == 0 and matte == 0)
    lib().mvo_alpha_composite(_ptr(bg), bg.shape[1], bg.shape[0], bg.strides[0], _ptr(fg), fg.shape[1],
                              fg.shape[0], fg.strides[0], int(position[0]), int(position[1]),
                              float(opacity), mode, matte, int(bool(use_pil)))
    return bg


def blur_ksize(radius: float) -> int:
    return lib().mvo_blur_ksize(float(radius))


def gaussian_weights_q8(sigma: float, k: int) -> np.ndarray:
    w = np.empty(k, dtype=np.int32)
    lib().mvo_gaussian_weights_q8(float(sigma), int(k), _ptr(w))
    return w


def gaussian_blur_u8(src: np.ndarray, k: int, sigma: float) -> np.ndarray:
    """cv2.GaussianBlur(src, (k, k), sigmaX=sigma, sigmaY=sigma) for uint8 (H,W) or (H,W,4)."""
    src = np.ascontiguousarray(src)
    assert src.dtype == np.uint8
    Cn = 1 if src.ndim == 2 else src.shape[2]
    assert Cn in (1, 4)
    w = gaussian_weights_q8(sigma, k)
    dst = np.empty_like(src)
    lib().mvo_gaussian_blur_u8(_ptr(src), src.shape[1], src.shape[0], src.strides[0], Cn, k, _ptr(w),
                               _ptr(dst), dst.strides[0])
    return dst


def effect_gaussian_blur(img: np.ndarray, radius: float) -> np.ndarray:
    """GaussianBlur.__call__ (effect/blur.py:29-43)."""
    img = _rgba(img)
    if radius == 0:
        return img
    k = blur_ksize(radius)
    out = np.empty((img.shape[0] + 2 * k, img.shape[1] + 2 * k, 4), dtype=np.uint8)
    lib().mvo_effect_gaussian_blur(_ptr(img), img.shape[1], img.shape[0], img.strides[0], float(radius),
                                   _ptr(out), out.strides[0])
    return out


def glow_strength_lut(strength: float) -> np.ndarray:
    """effect/blur.py:82: clip(strength * float32(v), 0, 255).astype(uint8), numpy-evaluated."""
    v = np.arange(256, dtype=np.uint8).astype(np.float32)
    return np.clip(float(strength) * v, 0, 255).astype(np.uint8)


def effect_glow(img: np.ndarray, radius: float, strength: float = 1.0) -> np.ndarray:
    """Glow.__call__ (effect/blur.py:66-85)."""
    img = _rgba(img)
    if radius == 0:
        return img
    k = blur_ksize(radius)
    out = np.empty((img.shape[0] + 2 * k, img.shape[1] + 2 * k, 4), dtype=np.uint8)
    lut = glow_strength_lut(strength)
    lib().mvo_effect_glow(_ptr(img), img.shape[1], img.shape[0], img.strides[0], float(radius), _ptr(lut),
                          _ptr(out), out.strides[0])
    return out


def shadow_opacity_lut(opacity: float) -> np.ndarray:
    """effect/style.py:64,68: (opacity[1] float64 * uint8 image).astype(uint8)."""
    return (np.array([float(opacity)]) * np.arange(256, dtype=np.uint8)).astype(np.uint8)


def effect_drop_shadow(img: np.ndarray, radius: float = 0.0, offset: float = 0.0, angle: float = 45.0,
                       color=(0, 0, 0), opacity: float = 0.5) -> np.ndarray:
    """DropShadow.__call__ (effect/style.py:49-84)."""
    img = _rgba(img)
    k = blur_ksize(radius) if radius > 0 else 0
    theta = float(angle) * np.pi / 180
    vec = np.round(float(offset) * np.array([np.cos(theta), np.sin(theta)])).astype(int)
    vx, vy = int(vec[0]), int(vec[1])
    out = np.empty((img.shape[0] + 2 * k + 2 * abs(vy), img.shape[1] + 2 * k + 2 * abs(vx), 4), dtype=np.uint8)
    col = np.round(np.asarray(color, dtype=np.float64)).astype(np.uint8)
    lut = shadow_opacity_lut(opacity)
    lib().mvo_effect_drop_shadow(_ptr(img), img.shape[1], img.shape[0], img.strides[0], float(radius),
                                 _ptr(lut), _ptr(col), vx, vy, _ptr(out), out.strides[0])
    return out


def rgb2hsv(rgb: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(rgb, COLOR_RGB2HSV) for uint8 (..., 3|4) -> (..., 3)."""
    rgb = np.ascontiguousarray(rgb)
    cn = rgb.shape[-1]
    n = rgb.size // cn
    out = np.empty(rgb.shape[:-1] + (3,), dtype=np.uint8)
    lib().mvo_rgb2hsv(_ptr(rgb), n, cn, _ptr(out))
    return out


def hsv2rgb(hsv: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(hsv, COLOR_HSV2RGB) for a uint8 (H, W, 3) image (row-tail quirk included)."""
    hsv = np.ascontiguousarray(hsv)
    assert hsv.ndim == 3 and hsv.shape[2] == 3
    out = np.empty_like(hsv)
    lib().mvo_hsv2rgb(_ptr(hsv), hsv.shape[1], hsv.shape[0], _ptr(out))
    return out


def chroma_key_bounds(key_color=(0, 255, 0), key_color_range=(20.0, 0.3, 0.3)):
    """ChromaKey.__init__ (contrib/segmentation.py:40-56)."""
    c = rgb2hsv(np.array(key_color, dtype=np.uint8).reshape(1, 1, 3))[0, 0].astype(np.float64)
    r = key_color_range
    lo = np.clip(np.array([c[0] - r[0] / 2, c[1] - 255 * r[1], c[2] - 255 * r[2]]), 0, 255).astype(np.uint8)
    hi = np.clip(np.array([c[0] + r[0] / 2, c[1] + 255 * r[1], c[2] + 255 * r[2]]), 0, 255).astype(np.uint8)
    return lo, hi


def effect_chroma_key(img: np.ndarray, key_color=(0, 255, 0), key_color_range=(20.0, 0.3, 0.3)) -> np.ndarray:
    """ChromaKey.__call__ (contrib/segmentation.py:58-64)."""
    img = _rgba(img)
    lo, hi = chroma_key_bounds(key_color, key_color_range)
    out = np.empty_like(img)
    lib().mvo_effect_chroma_key(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(lo), _ptr(hi),
                                _ptr(out), out.strides[0])
    return out


def effect_fill_color(img: np.ndarray, color) -> np.ndarray:
    """FillColor.__call__ (effect/color.py:25-31)."""
    img = _rgba(img)
    col = np.round(np.asarray(color, dtype=np.float64)).astype(np.uint8)
    out = np.empty_like(img)
    lib().mvo_effect_fill_color(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(col), _ptr(out),
                                out.strides[0])
    return out


def hsl_shift_luts(hue: float, saturation: float, luminance: float):
    """effect/color.py:61-66 evaluated on every possible uint8 input (float32, numpy's own rounding)."""
    x = np.arange(256, dtype=np.uint8).astype(np.float32)
    dh, ds, dl = np.float32(hue), np.float32(saturation), np.float32(luminance)
    h = np.round((x + dh / 2) % 180).astype(np.uint8)
    s = np.clip(np.round(x + ds * 255), 0, 255).astype(np.uint8)
    v = np.clip(np.round(x + dl * 255), 0, 255).astype(np.uint8)
    return h, s, v


def effect_hsl_shift(img: np.ndarray, hue: float = 0.0, saturation: float = 0.0, luminance: float = 0.0) -> np.ndarray:
    """HSLShift.__call__ (effect/color.py:56-69)."""
    img = _rgba(img)
    h, s, v = hsl_shift_luts(hue, saturation, luminance)
    out = np.empty_like(img)
    lib().mvo_effect_hsl_shift(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(h), _ptr(s), _ptr(v),
                               _ptr(out), out.strides[0])
    return out


def stripe(size: tuple[int, int], angle: float = 45.0, color1=(0, 0, 0), color2=(255, 255, 255),
           opacity1: float = 1.0, opacity2: float = 1.0, total_width: float = 64.0, phase: float = 0.0,
           ratio: float = 0.5) -> np.ndarray:
    """Stripe.__call__ (layer/texture.py:159-191) for already-evaluated attribute values."""
    width, height = size
    c1 = np.array([*np.round(np.asarray(color1, dtype=np.float64)), np.round(255 * float(opacity1))], dtype=np.float64)
    c2 = np.array([*np.round(np.asarray(color2, dtype=np.float64)), np.round(255 * float(opacity2))], dtype=np.float64)
    ratio = float(ratio)
    solid = 1 if ratio <= 0.0 else (2 if ratio >= 1.0 else 0)
    theta = float(angle) / 180.0 * np.pi
    v = np.array([np.sin(theta), np.cos(theta)], dtype=np.float64)
    eps = 1e-2
    out = np.empty((height, width, 4), dtype=np.uint8)
    lib().mvo_stripe(_ptr(out), width, height, out.strides[0], _ptr(v), float(total_width), float(phase),
                     max(ratio - eps, 0.0), min(ratio + eps, 1.0), _ptr(c1), _ptr(c2), solid)
    return out


# ---------------------------------------------------------------------------------------------
# host geometry (numpy float64, same operations in the same order as the reference)
# ---------------------------------------------------------------------------------------------

_ORIGIN_VECTORS = {  # enum.py:267-288, keyed by Direction value / name
    "bottom_left": (0.0, 1.0), "bottom_center": (0.5, 1.0), "bottom_right": (1.0, 1.0),
    "center_left": (0.0, 0.5), "center": (0.5, 0.5), "center_right": (1.0, 0.5),
    "top_left": (0.0, 0.0), "top_center": (0.5, 0.0), "top_right": (1.0, 0.0),
}


@dataclass
class TransformSample:
    """Plain-number stand-in for movis.transform.TransformValue (transform.py:12-42)."""
    anchor_point: tuple[float, float] = (0.0, 0.0)
    position: tuple[float, float] = (0.0, 0.0)
    scale: tuple[float, float] = (1.0, 1.0)
    rotation: float = 0.0
    opacity: float = 1.0
    origin_point: str = "center"
    blending_mode: str = "normal"


def origin_vector(origin: str, size: tuple[float, float]) -> tuple[float, float]:
    """Direction.to_vector (enum.py:267-288): size/2 where the reference halves, size or 0 otherwise."""
    fx, fy = _ORIGIN_VECTORS[origin]
    vx = 0 if fx == 0.0 else (size[0] / 2 if fx == 0.5 else size[0])
    vy = 0 if fy == 0.0 else (size[1] / 2 if fy == 0.5 else size[1])
    return vx, vy


def layer_matrix(p: TransformSample, size: tuple[int, int]) -> np.ndarray:
    """M = T1 @ SR @ T2 (composition.py:879-881, :910-934), 3x3 float64."""
    w, h = float(size[0]), float(size[1])
    T1 = np.array([[1, 0, p.position[0] + p.anchor_point[0]],
                   [0, 1, p.position[1] + p.anchor_point[1]],
                   [0, 0, 1]], dtype=np.float64)
    cos_t = np.cos((2 * np.pi * p.rotation) / 360)
    sin_t = np.sin((2 * np.pi * p.rotation) / 360)
    SR = np.array([[p.scale[0] * cos_t, -p.scale[0] * sin_t, 0],
                   [p.scale[1] * sin_t, p.scale[1] * cos_t, 0],
                   [0, 0, 1]], dtype=np.float64)
    c = origin_vector(p.origin_point, (w, h))
    T2 = np.array([[1, 0, -p.anchor_point[0] - c[0]],
                   [0, 1, -p.anchor_point[1] - c[1]],
                   [0, 0, 1]], dtype=np.float64)
    return T1 @ SR @ T2


def fixed_affine(p: TransformSample, size: tuple[int, int], preview_level: int = 1):
    """_get_fixed_affine_matrix (composition.py:873-907): (A 2x3, (W, H), (ox, oy)) or None."""
    w, h = size
    M = layer_matrix(p, size)
    L = preview_level
    P = np.array([[1 / L, 0, 0], [0, 1 / L, 0], [0, 0, 1]], dtype=np.float64)
    A = (P @ M)[:2]
    corners = np.array([[0, 0, 1], [0, h, 1], [w, 0, 1], [w, h, 1]], dtype=np.float64)
    cg = corners @ A.transpose()
    mn = np.ceil(cg.min(axis=0))
    mx = np.floor(cg.max(axis=0))
    WH = (mx - mn).astype(np.int32)
    W, H = int(WH[0]), int(WH[1])
    if W == 0 or H == 0:
        return None
    ox, oy = int(mn[0]), int(mn[1])
    Pf = np.array([[1 / L, 0, -ox], [0, 1 / L, -oy], [0, 0, 1]], dtype=np.float64)
    return (Pf @ M)[:2], (W, H), (ox, oy)


@dataclass
class LayerPlan:
    """One visible layer of one frame: post-effect image + evaluated transform."""
    image: np.ndarray
    transform: TransformSample


def composite_frame(size: tuple[int, int], layers: Sequence[LayerPlan], bg_color=(0, 0, 0, 0),
                    preview_level: int = 1) -> np.ndarray:
    """Composition.__call__ for already-evaluated layers (composition.py:345-370, :791-820)."""
    L = preview_level
    H, W = size[1] // L, size[0] // L
    frame = np.empty((H, W, 4), dtype=np.uint8)
    descs = (_Layer * max(1, len(layers)))()
    keep = []
    n = 0
    for lp in layers:
        img = _rgba(lp.image)
        keep.append(img)
        r = fixed_affine(lp.transform, (img.shape[1], img.shape[0]), L)
        if r is None:
            continue
        A, (bw, bh), (ox, oy) = r
        if bw < 0 or bh < 0:
            # cv2 treats a negative dsize as "use the source size" (SURVEY.md App. B); both this
            # oracle and the product skip such degenerate layers (documented divergence).
            continue
        d = descs[n]
        d.src, d.src_w, d.src_h, d.src_stride = _ptr(img), img.shape[1], img.shape[0], img.strides[0]
        for j, v in enumerate(np.asarray(A, dtype=np.float64).reshape(6)):
            d.M[j] = float(v)
        d.bbox_x, d.bbox_y, d.bbox_w, d.bbox_h = ox, oy, bw, bh
        mode = _mode_index(lp.transform.blending_mode)
        d.blend_mode, d.use_pil, d.opacity = mode, int(mode == 0), float(lp.transform.opacity)
        n += 1
    bgc = np.asarray(bg_color, dtype=np.uint8)
    rc = lib().mvo_composite_stack(_ptr(frame), W, H, frame.strides[0], _ptr(bgc), n, descs)
    if rc != 0:
        raise MemoryError("oracle composite_stack failed")
    return frame

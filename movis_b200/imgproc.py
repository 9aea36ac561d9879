"""``alpha_composite`` -- the public blend operator of the render path, on the GPU.

Drop-in for ``movis.imgproc.alpha_composite`` (reference: imgproc.py:216-273): same signature,
argument meaning, asserts and dispatch quirks; the pixels are computed by the sm_100a kernel
behind ``mvb_alpha_composite`` and are bit-identical to the reference's PIL / numpy results.

Inputs may be numpy arrays (uploaded, result downloaded -- API compatibility) or CUDA uint8
tensors (stay on the device; this is what the effects and matte layers use internally).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native
from .device import image_args, require_cuda, stream_ptr, to_device
from .enum import BlendingMode, MatteMode


def _resolve(blending_mode, matte_mode) -> tuple[int, int, int]:
    """(blend code, matte code, flags) following the reference's dispatch (imgproc.py:265-273)."""
    if blending_mode == BlendingMode.NORMAL and matte_mode == MatteMode.NONE:
        return BlendingMode.NORMAL.value, MatteMode.NONE.value, _native.MVB_LAYER_PIL_OVER
    mode = BlendingMode.from_string(blending_mode) if isinstance(blending_mode, str) else blending_mode
    if isinstance(matte_mode, MatteMode):
        matte = matte_mode.value
    else:
        # The reference compares matte_mode against the enum only (imgproc.py:147,166-169): any
        # other object (e.g. the string 'none') matches no branch, so alpha is left untouched --
        # the same observable behaviour as MatteMode.ALPHA.
        matte = MatteMode.ALPHA.value
    return mode.value, matte, 0


def composite_device(bg: torch.Tensor, fg: torch.Tensor, position=(0, 0), opacity: float = 1.0,
                     blend: int = 0, matte: int = 0, flags: int = 0) -> torch.Tensor:
    """In-place composite of two CUDA images through the C ABI; returns ``bg``."""
    bp, bw, bh, bs = image_args(bg)
    fp, fw, fh, fs = image_args(fg)
    _native.check(_native.lib().mvb_alpha_composite(
        bp, bw, bh, bs, fp, fw, fh, fs, int(position[0]), int(position[1]), float(opacity),
        int(blend), int(matte), int(flags), stream_ptr()))
    return bg


def alpha_composite(bg_image, fg_image, position: tuple[int, int] = (0, 0), opacity: float = 1.0,
                    blending_mode: str | BlendingMode = BlendingMode.NORMAL,
                    matte_mode: MatteMode = MatteMode.NONE):
    """Alpha-composite ``fg_image`` over ``bg_image`` at ``position``.

    Returns the composited image (same kind as ``bg_image``: numpy array or CUDA tensor).  As in
    the reference, callers must use the return value: NORMAL blending without matte returns a new
    image and leaves ``bg_image`` untouched, every other combination updates ``bg_image`` in
    place (a read-only numpy background is copied first).
    """
    for name, img in (("bg_image", bg_image), ("fg_image", fg_image)):
        assert img.ndim == 3, f"{name} must have 3 dimensions"
        assert img.shape[2] == 4, f"{name} must have 4 channels (RGBA)"
        assert img.dtype in (np.uint8, torch.uint8), f"{name} must be uint8"
    blend, matte, flags = _resolve(blending_mode, matte_mode)
    use_pil = bool(flags & _native.MVB_LAYER_PIL_OVER)
    if use_pil:
        assert 0.0 <= opacity <= 1.0, f"opacity must be in [0, 1], but {opacity} is given."
    opacity = min(float(opacity), 1.0)  # imgproc.py:149,159: opacity >= 1 means "unscaled"
    require_cuda()

    host_bg = isinstance(bg_image, np.ndarray)
    fg_dev = to_device(fg_image, "fg_image")
    bg_dev = to_device(bg_image, "bg_image")  # a private upload for numpy input
    if not host_bg and use_pil and bg_dev is bg_image:
        bg_dev = bg_dev.clone()  # the PIL path never mutates its background (imgproc.py:210-213)
    composite_device(bg_dev, fg_dev, position, opacity, blend, matte, flags)
    if not host_bg:
        if not use_pil and bg_dev is not bg_image and bg_image.is_cuda:
            bg_image.copy_(bg_dev)
            return bg_image
        return bg_dev
    result = bg_dev.cpu().numpy()
    if use_pil:
        return result
    if not bg_image.flags.writeable:
        return result
    bg_image[...] = result
    return bg_image

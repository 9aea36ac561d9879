"""GaussianBlur and Glow on the GPU (reference: movis/effect/blur.py).

Both pad the layer by ``k = 2*ceil(max(1, (4r-1)/2)) + 1`` px per side and run cv2's bit-exact
fixed-point Gaussian; here pad + separable blur (+ the Glow composite) are one C-ABI call,
``mvb_effect_blur_rgba8``.  The float32 strength scaling of Glow is handed to the kernel as a
256-entry table evaluated with numpy, so its rounding is numpy's own (blur.py:82).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..attribute import Attribute, AttributesMixin, AttributeType
from ..device import empty_image, image_args, stream_ptr
from .protocol import DeviceEffectMixin


def _blur_device(image: torch.Tensor, radius: float, mode: int, strength_lut: np.ndarray | None) -> torch.Tensor:
    L = _native.lib()
    k = _native.blur_ksize(radius)
    weights = _native.gaussian_weights_q8(radius, k)
    ptr, w, h, stride = image_args(image)
    out = empty_image(h + 2 * k, w + 2 * k)
    scratch = torch.empty(L.mvb_effect_blur_scratch_bytes(w, h, k, 4), dtype=torch.uint8, device="cuda")
    lut_ptr = strength_lut.ctypes.data if strength_lut is not None else None
    _native.check(L.mvb_effect_blur_rgba8(ptr, w, h, stride, out.data_ptr(), out.stride(0), k,
                                          weights.ctypes.data, mode, lut_ptr, scratch.data_ptr(),
                                          stream_ptr()))
    return out


class GaussianBlur(DeviceEffectMixin, AttributesMixin):
    """Gaussian blur; the output is ``2k`` px larger in each dimension so blurred edges survive.

    Animatable attribute: ``radius``."""

    def __init__(self, radius: float):
        self.radius = Attribute(radius, AttributeType.SCALAR, range=(0., 1e6))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        radius = float(self.radius(time)[0])
        if radius == 0:
            return image
        assert 0 < radius, f"radius must be nonnegative, but {radius}"
        return _blur_device(image, radius, _native.MVB_BLUR_PLAIN, None)


class Glow(DeviceEffectMixin, AttributesMixin):
    """Adds a ``strength``-scaled Gaussian blur of the layer onto itself (LINEAR_DODGE).

    Animatable attributes: ``radius``, ``strength``."""

    def __init__(self, radius: float, strength: float = 1.0):
        self.radius = Attribute(radius, AttributeType.SCALAR, range=(0., 1e6))
        self.strength = Attribute(strength, AttributeType.SCALAR, range=(0., 100.))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        radius = float(self.radius(time)[0])
        if radius == 0.:
            return image
        assert 0 < radius, f"radius must be positive, but {radius}"
        strength = float(self.strength(time)[0])
        levels = np.arange(256, dtype=np.uint8).astype(np.float32)
        lut = np.clip(strength * levels, 0, 255).astype(np.uint8)  # blur.py:82 on every uint8 value
        return _blur_device(image, radius, _native.MVB_BLUR_GLOW, lut)

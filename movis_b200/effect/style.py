"""DropShadow on the GPU (reference: movis/effect/style.py)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..attribute import Attribute, AttributesMixin, AttributeType
from ..device import empty_image, image_args, stream_ptr
from ..util import to_rgb
from .protocol import DeviceEffectMixin


class DropShadow(DeviceEffectMixin, AttributesMixin):
    """Blurred, offset, tinted copy of the layer's alpha placed behind the layer.

    Animatable attributes: ``radius``, ``offset``, ``angle``, ``color``, ``opacity``."""

    def __init__(self, radius: float = 0.0, offset: float = 0.0, angle: float = 45.0,
                 color: tuple[int, int, int] | str = (0, 0, 0), opacity: float = 0.5) -> None:
        self.radius = Attribute(radius, AttributeType.SCALAR, range=(0., 1e6))
        self.angle = Attribute(angle, AttributeType.SCALAR)
        self.offset = Attribute(offset, AttributeType.SCALAR)
        self.color = Attribute(to_rgb(color), AttributeType.COLOR, range=(0., 255.))
        self.opacity = Attribute(opacity, AttributeType.SCALAR, range=(0., 1.))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        L = _native.lib()
        radius = float(self.radius(time)[0])
        assert 0 <= radius, f"radius must be nonnegative, but {radius}"
        k = _native.blur_ksize(radius) if radius > 0 else 0
        weights = _native.gaussian_weights_q8(radius, k) if k else None
        # style.py:64,68: (opacity * uint8 image).astype(uint8), float64 -- tabulated for all 256 inputs
        lut = (self.opacity(time) * np.arange(256, dtype=np.uint8)).astype(np.uint8)
        color = np.round(self.color(time)).astype(np.uint8)
        theta = float(self.angle(time)[0]) * np.pi / 180
        vec = np.round(float(self.offset(time)[0]) * np.array([np.cos(theta), np.sin(theta)])).astype(int)
        vx, vy = int(vec[0]), int(vec[1])
        ptr, w, h, stride = image_args(image)
        out = empty_image(h + 2 * k + 2 * abs(vy), w + 2 * k + 2 * abs(vx))
        scratch = None
        if k:
            n = L.mvb_effect_blur_scratch_bytes(w, h, k, 1) + (w + 2 * k) * (h + 2 * k)
            scratch = torch.empty(n, dtype=torch.uint8, device="cuda")
        _native.check(L.mvb_effect_drop_shadow_rgba8(
            ptr, w, h, stride, out.data_ptr(), out.stride(0), k,
            weights.ctypes.data if weights is not None else None, lut.ctypes.data, color.ctypes.data,
            vx, vy, scratch.data_ptr() if scratch is not None else None, stream_ptr()))
        return out

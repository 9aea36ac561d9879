"""FillColor and HSLShift on the GPU (reference: movis/effect/color.py)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..attribute import Attribute, AttributesMixin, AttributeType
from ..device import empty_image, image_args, stream_ptr
from ..util import to_rgb
from .protocol import DeviceEffectMixin


class FillColor(DeviceEffectMixin, AttributesMixin):
    """Replace RGB by a colour, keep alpha.  Animatable attribute: ``color``."""

    def __init__(self, color: tuple[int, int, int] | str = (255, 255, 255)):
        self.color = Attribute(to_rgb(color), AttributeType.COLOR, range=(0., 255.))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        color = np.round(self.color(time)).astype(np.uint8)
        ptr, w, h, stride = image_args(image)
        out = empty_image(h, w)
        _native.check(_native.lib().mvb_effect_fill_color_rgba8(
            ptr, w, h, stride, out.data_ptr(), out.stride(0), color.ctypes.data, stream_ptr()))
        return out


class HSLShift(DeviceEffectMixin, AttributesMixin):
    """Shift hue (degrees), saturation and luminance ([-1, 1]) in cv2's 8-bit HSV space.

    Animatable attributes: ``hue``, ``saturation``, ``luminance``."""

    def __init__(self, hue: float = 0.0, saturation: float = 0.0, luminance: float = 0.0):
        self.hue = Attribute(hue, AttributeType.SCALAR)
        self.saturation = Attribute(saturation, AttributeType.SCALAR, range=(-1., 1.))
        self.luminance = Attribute(luminance, AttributeType.SCALAR, range=(-1., 1.))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        # color.py:60-66 applied to every possible uint8 channel value (float32, numpy rounding)
        x = np.arange(256, dtype=np.uint8).astype(np.float32)
        dh = np.float32(self.hue(time))
        ds = np.float32(self.saturation(time))
        dl = np.float32(self.luminance(time))
        h_lut = np.round((x + dh / 2) % 180).astype(np.uint8)
        s_lut = np.clip(np.round(x + ds * 255), 0, 255).astype(np.uint8)
        v_lut = np.clip(np.round(x + dl * 255), 0, 255).astype(np.uint8)
        ptr, w, h, stride = image_args(image)
        out = empty_image(h, w)
        _native.check(_native.lib().mvb_effect_hsl_shift_rgba8(
            ptr, w, h, stride, out.data_ptr(), out.stride(0), h_lut.ctypes.data, s_lut.ctypes.data,
            v_lut.ctypes.data, stream_ptr()))
        return out

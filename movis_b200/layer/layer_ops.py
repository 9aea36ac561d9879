"""AlphaMatte / LuminanceMatte layers on the GPU (reference: movis/layer/layer_ops.py)."""
from __future__ import annotations

from typing import Hashable

import numpy as np

from ..attribute import Attribute, AttributesMixin, AttributeType
from ..device import check_image_array, to_device, to_host
from ..enum import BlendingMode, MatteMode
from ..imgproc import composite_device


def _sub_device(layer, time: float):
    if hasattr(layer, "render_device"):
        return layer.render_device(time)
    host = layer(time)
    if host is None:
        return None
    check_image_array(host, "Rendered layer image")
    return to_device(host)


def _sub_key(layer, time: float) -> Hashable:
    return layer.get_key(time) if hasattr(layer, "get_key") else time


class AlphaMatte(AttributesMixin):
    """Overlay ``target`` on ``mask`` without changing the mask's alpha channel.

    Animatable attribute: ``opacity``."""

    def __init__(self, mask, target, opacity: float = 1.0, blending_mode: BlendingMode | str = BlendingMode.NORMAL):
        self.mask = mask
        self.target = target
        self.opacity = Attribute(opacity, value_type=AttributeType.SCALAR, range=(0., 1.))
        self.blending_mode = BlendingMode.from_string(blending_mode) \
            if isinstance(blending_mode, str) else blending_mode

    def get_key(self, time: float) -> tuple[Hashable, Hashable, Hashable]:
        return (super().get_key(time), _sub_key(self.mask, time), _sub_key(self.target, time))

    @property
    def duration(self) -> float:
        return self.mask.duration

    def render_device(self, time: float):
        if time < 0 or self.duration <= time:
            return None
        mask = _sub_device(self.mask, time)
        if mask is None:
            return None
        target = _sub_device(self.target, time)
        if target is None:
            return mask
        opacity = float(self.opacity(time)[0])
        return composite_device(mask.clone(), target, (0, 0), opacity, self.blending_mode.value,
                                MatteMode.ALPHA.value, 0)

    def __call__(self, time: float) -> np.ndarray | None:
        frame = self.render_device(time)
        return None if frame is None else to_host(frame)


class LuminanceMatte:
    """Replace the alpha of ``target`` by the luminance of ``mask``."""

    def __init__(self, mask, target):
        self.mask = mask
        self.target = target

    def get_key(self, time: float) -> tuple[Hashable, Hashable]:
        return (_sub_key(self.mask, time), _sub_key(self.target, time))

    @property
    def duration(self) -> float:
        return self.mask.duration

    def render_device(self, time: float):
        if time < 0 or self.duration <= time:
            return None
        mask = _sub_device(self.mask, time)
        if mask is None:
            return None
        target = _sub_device(self.target, time)
        if target is None:
            return None
        return composite_device(mask.clone(), target, (0, 0), 1.0, BlendingMode.NORMAL.value,
                                MatteMode.LUMINANCE.value, 0)

    def __call__(self, time: float) -> np.ndarray | None:
        frame = self.render_device(time)
        return None if frame is None else to_host(frame)

"""Effect protocol (reference: movis/effect/protocol.py).

An effect is any callable ``(prev_image, time) -> image`` on uint8 RGBA ``ndarray``s; the output
may have a different size.  ``get_key(time)`` is optional.  Effects of this package additionally
implement ``apply_device(image, time)`` on CUDA tensors, which the compositor prefers.
"""
from __future__ import annotations

from typing import Hashable, Protocol

import numpy as np


class Effect(Protocol):
    def __call__(self, prev_image: np.ndarray, time: float) -> np.ndarray:
        raise NotImplementedError


class BasicEffect(Effect):
    def get_key(self, time: float) -> Hashable:
        """Hashable state at ``time``; effects without it are assumed to change every frame."""
        return time


class DeviceEffectMixin:
    """Gives a device effect the numpy-facing ``__call__`` of the protocol."""

    def __call__(self, prev_image, time: float):
        from ..device import is_device_image, to_device, to_host
        if is_device_image(prev_image):
            return self.apply_device(to_device(prev_image), time)
        assert prev_image.ndim == 3
        assert prev_image.shape[2] == 4, f"prev_image must be RGBA image, but {prev_image.shape}"
        return to_host(self.apply_device(to_device(prev_image), time))

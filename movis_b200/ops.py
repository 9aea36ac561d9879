"""``crop`` on the GPU (reference: movis/ops.py:280-321).

The other helpers of the reference's ``ops`` module (concatenate, repeat, trim, tile, ...) are
timeline builders made of public ``Composition`` calls; they are outside the per-frame render path
this package replaces and work unchanged on top of ``movis_b200.layer.Composition``.
"""
from __future__ import annotations

from typing import Hashable

import numpy as np

from .device import to_host
from .layer.layer_ops import _sub_device


class _CropLayer:
    """``layer`` restricted to ``rect = (x, y, width, height)`` (ops.py:280-305).

    On the device the crop is a strided VIEW of the wrapped layer's image: no pixels move, and the
    fused kernel samples it through its row stride like any other source (the reference's numpy view
    ``img[y:y+h, x:x+w]`` handed to cv2.warpAffine, ops.py:296)."""

    def __init__(self, layer, rect: tuple[int, int, int, int]):
        assert len(rect) == 4
        self.layer = layer
        self.rect = rect

    @property
    def duration(self) -> float:
        return self.layer.duration

    def render_device(self, time: float):
        img = _sub_device(self.layer, time)
        if img is None:
            return None
        x, y, w, h = self.rect
        return img[y:y + h, x:x + w]   # same slice semantics as numpy (clamped to the image)

    def __call__(self, time: float) -> np.ndarray | None:
        frame = self.render_device(time)
        return None if frame is None else to_host(frame)

    def get_key(self, time: float) -> Hashable:
        return self.layer.get_key(time)

    def get_audio(self, start_time: float, end_time: float) -> np.ndarray | None:
        if hasattr(self.layer, "get_audio"):
            return self.layer.get_audio(start_time, end_time)
        return None


def crop(layer, rect: tuple[int, int, int, int]) -> _CropLayer:
    """Crop a layer to ``rect = (x, y, width, height)``; see ``movis.ops.crop``."""
    return _CropLayer(layer, rect)

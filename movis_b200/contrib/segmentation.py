"""ChromaKey on the GPU (reference: movis/contrib/segmentation.py:21-64).

The only colour key in the reference.  ``RobustVideoMatting`` (an ONNX model) is out of scope.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..device import attach_alpha_map, empty_image, image_args, stream_ptr
from ..effect.protocol import DeviceEffectMixin
from ..util import to_rgb


class ChromaKey(DeviceEffectMixin):
    """Make pixels whose cv2-HSV colour lies within ``key_color_range`` of ``key_color``
    transparent; every other pixel becomes fully opaque (the input alpha is discarded).

    Args:
        key_color: RGB tuple or CSS colour name.
        color_space: only ``'hsv'`` is supported.
        key_color_range: full hue width in degrees, saturation and value half-widths in [0, 1].
    """

    def __init__(self, key_color: tuple[int, int, int] | str = (0, 255, 0), color_space: str = "hsv",
                 key_color_range: tuple[float, float, float] = (20.0, 0.3, 0.3)):
        assert color_space == "hsv", "Only HSV color space is supported."
        c = _native.rgb_to_hsv_u8(to_rgb(key_color)).astype(np.float64)
        spread = np.array([key_color_range[0] / 2, 255 * key_color_range[1], 255 * key_color_range[2]])
        self._key_color = c.astype(np.uint8)
        self._lower_color = np.clip(c - spread, 0, 255).astype(np.uint8)
        self._upper_color = np.clip(c + spread, 0, 255).astype(np.uint8)

    def get_key(self, time: float) -> tuple:
        """The keyed pixels depend on the input image and the constructor arguments only, so equal
        source frames may share one cached result (the reference's ChromaKey defines no ``get_key`` and its
        layer cache therefore never hits, composition.py:767-769)."""
        return (tuple(int(v) for v in self._lower_color), tuple(int(v) for v in self._upper_color))

    def apply_device(self, image: torch.Tensor, time: float) -> torch.Tensor:
        ptr, w, h, stride = image_args(image)
        out = empty_image(h, w)
        _native.check(_native.lib().mvb_effect_chroma_key_rgba8(
            ptr, w, h, stride, out.data_ptr(), out.stride(0), self._lower_color.ctypes.data,
            self._upper_color.ctypes.data, stream_ptr()))
        # keyed-out regions are what a chroma key is for: let the compositor skip them tile by tile
        return attach_alpha_map(out)

"""Procedural texture sources (reference: movis/layer/texture.py).

``Stripe`` is pure float64 math in the reference (texture.py:159-191), so it is rendered on the
device with the same IEEE operations in the same order (``mvb_source_stripe_rgba8``) and never
touches the host.  ``Gradient`` paints through Qt (QLinearGradient / QRadialGradient) and is not
part of this package.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native
from ..attribute import Attribute, AttributesMixin, AttributeType
from ..device import require_cuda, stream_ptr, to_host
from ..util import to_rgb


class Stripe(AttributesMixin):
    """A layer that generates a stripe pattern; same arguments and animatable attributes as the
    reference (``angle``, ``color1``, ``color2``, ``opacity1``, ``opacity2``, ``total_width``, ``phase``,
    ``ratio``)."""

    def __init__(
        self,
        size: tuple[int, int] = (100, 100),
        angle: float = 45.0,
        color1: tuple[int, int, int] | str = (0, 0, 0),
        color2: tuple[int, int, int] | str = (255, 255, 255),
        opacity1: float = 1.0,
        opacity2: float = 1.0,
        total_width: float = 64.0,
        phase: float = 0.0,
        ratio: float = 0.5,
        duration: float = 1e6,
    ) -> None:
        self.size = size
        self.duration = duration
        self.angle = Attribute(angle, AttributeType.ANGLE)
        self.color1 = Attribute(to_rgb(color1), AttributeType.COLOR, range=(0.0, 255.0))
        self.color2 = Attribute(to_rgb(color2), AttributeType.COLOR, range=(0.0, 255.0))
        self.opacity1 = Attribute(opacity1, AttributeType.SCALAR, range=(0.0, 1.0))
        self.opacity2 = Attribute(opacity2, AttributeType.SCALAR, range=(0.0, 1.0))
        self.total_width = Attribute(total_width, AttributeType.SCALAR, range=(0.0, 1e6))
        self.phase = Attribute(phase, AttributeType.SCALAR)
        self.ratio = Attribute(ratio, AttributeType.SCALAR, range=(0.0, 1.0))

    def stripe_params(self, time: float) -> tuple[np.ndarray, int]:
        """The 16 doubles of ``mvb_source_stripe_rgba8`` and its ``solid`` flag, evaluated exactly as
        texture.py:162-186 evaluates them on the host."""
        width, height = self.size
        ratio = float(self.ratio(time)[0])
        c1 = np.concatenate([np.round(self.color1(time)).reshape(3), [np.round(255 * float(self.opacity1(time)[0]))]])
        c2 = np.concatenate([np.round(self.color2(time)).reshape(3), [np.round(255 * float(self.opacity2(time)[0]))]])
        solid = 1 if ratio <= 0.0 else (2 if ratio >= 1.0 else 0)
        theta = float(self.angle(time)[0]) / 180.0 * np.pi
        eps = 1e-2
        edge0, edge1 = max(ratio - eps, 0.0), min(ratio + eps, 1.0)
        v = np.array([np.sin(theta), np.cos(theta)], dtype=np.float64)
        params = np.array([v[0], v[1], height / 2, width / 2, float(self.total_width(time)[0]), float(self.phase(time)[0]),
                           edge0, edge1 - edge0, *c1, *c2], dtype=np.float64)
        return params, solid

    def render_device(self, time: float):
        if time < 0 or time >= self.duration:
            return None
        require_cuda()
        width, height = self.size
        params, solid = self.stripe_params(time)
        out = torch.empty((height, width, 4), dtype=torch.uint8, device="cuda")
        _native.check(_native.lib().mvb_source_stripe_rgba8(
            out.data_ptr(), width, height, out.stride(0), params.ctypes.data, solid, stream_ptr()))
        return out

    def __call__(self, time: float) -> np.ndarray | None:
        image = self.render_device(time)
        return None if image is None else to_host(image)

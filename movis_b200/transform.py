"""Layer transform (host side of the render path).

API-compatible with the reference's ``movis/transform.py``: ``Transform`` groups the five
animatable attributes plus origin point and blending mode; ``get_current_value`` returns a
``TransformValue``.  ``Transform.sample`` evaluates a frame range at once for the batched planner.
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np

from .attribute import Attribute
from .enum import AttributeType, BlendingMode, Direction


class TransformValue(NamedTuple):
    """Evaluated transform of a layer at one instant (transform.py:12-42)."""
    anchor_point: tuple[float, float] = (0.0, 0.0)
    position: tuple[float, float] = (0.0, 0.0)
    scale: tuple[float, float] = (1.0, 1.0)
    rotation: float = 0.0
    opacity: float = 1.0
    origin_point: Direction = Direction.CENTER
    blending_mode: BlendingMode = BlendingMode.NORMAL


class TransformSamples(NamedTuple):
    """``Transform`` evaluated at N layer times (arrays are (N, 2) or (N,))."""
    anchor_point: np.ndarray
    position: np.ndarray
    scale: np.ndarray
    rotation: np.ndarray
    opacity: np.ndarray
    origin_point: Direction
    blending_mode: BlendingMode

    def at(self, i: int) -> TransformValue:
        return TransformValue(
            anchor_point=(float(self.anchor_point[i, 0]), float(self.anchor_point[i, 1])),
            position=(float(self.position[i, 0]), float(self.position[i, 1])),
            scale=(float(self.scale[i, 0]), float(self.scale[i, 1])),
            rotation=float(self.rotation[i]), opacity=float(self.opacity[i]),
            origin_point=self.origin_point, blending_mode=self.blending_mode)


_EDGE_ORIGINS = {
    # (top, bottom, left, right) given? -> origin
    (True, False, False, False): Direction.TOP_CENTER,
    (False, True, False, False): Direction.BOTTOM_CENTER,
    (False, False, True, False): Direction.CENTER_LEFT,
    (False, False, False, True): Direction.CENTER_RIGHT,
    (True, False, True, False): Direction.TOP_LEFT,
    (True, False, False, True): Direction.TOP_RIGHT,
    (False, True, True, False): Direction.BOTTOM_LEFT,
    (False, True, False, True): Direction.BOTTOM_RIGHT,
}


class Transform:
    """position / scale / rotation / opacity / anchor_point attributes + origin + blend mode
    (transform.py:45-101)."""

    def __init__(self, position=(0.0, 0.0), scale=(1.0, 1.0), rotation: float = 0.0, opacity: float = 1.0,
                 anchor_point=(0.0, 0.0), origin_point: Direction | str = Direction.CENTER,
                 blending_mode: BlendingMode | str = BlendingMode.NORMAL):
        self.position = Attribute(position, AttributeType.VECTOR2D)
        self.scale = Attribute(scale, AttributeType.VECTOR2D)
        self.rotation = Attribute(rotation, AttributeType.SCALAR)
        self.opacity = Attribute(opacity, AttributeType.SCALAR, range=(0., 1.))
        self.anchor_point = Attribute(anchor_point, AttributeType.VECTOR2D)
        self.origin_point = Direction.from_string(origin_point) if isinstance(origin_point, str) else origin_point
        self.blending_mode = BlendingMode.from_string(blending_mode) \
            if isinstance(blending_mode, str) else blending_mode

    @property
    def attributes(self) -> dict[str, Attribute]:
        return {"anchor_point": self.anchor_point, "position": self.position, "scale": self.scale,
                "rotation": self.rotation, "opacity": self.opacity}

    @property
    def is_static(self) -> bool:
        return all(a.is_static for a in self.attributes.values())

    @classmethod
    def from_positions(cls, size: tuple[int, int], top: float | None = None, bottom: float | None = None,
                       left: float | None = None, right: float | None = None,
                       object_fit: str | None = None) -> "Transform":
        """Place a layer by its distance to the canvas edges (transform.py:103-175)."""
        W, H = size[0], size[1]
        given = (top is not None, bottom is not None, left is not None, right is not None)
        if not any(given):
            if object_fit is None:
                scale = 1.0
            elif object_fit == "contain":
                scale = min(W, H) / max(W, H)
            elif object_fit == "cover":
                scale = max(W, H) / min(W, H)
            else:
                raise ValueError(f"Invalid object_fit: {object_fit}. contain or cover is expected.")
            return cls(position=(W / 2, H / 2), origin_point=Direction.CENTER, scale=scale)
        origin = _EDGE_ORIGINS.get(given)
        if origin is None:
            raise ValueError("Invalid pair of arguments")
        x = left if left is not None else (W - right if right is not None else W / 2)
        y = top if top is not None else (H - bottom if bottom is not None else H / 2)
        return cls(position=(x, y), origin_point=origin, scale=1.0)

    def get_current_value(self, layer_time: float) -> TransformValue:
        return TransformValue(
            anchor_point=transform_to_2dvector(self.anchor_point(layer_time)),
            position=transform_to_2dvector(self.position(layer_time)),
            scale=transform_to_2dvector(self.scale(layer_time)),
            rotation=transform_to_1dscalar(self.rotation(layer_time)),
            opacity=transform_to_1dscalar(self.opacity(layer_time)),
            origin_point=self.origin_point,
            blending_mode=self.blending_mode,
        )

    def sample(self, layer_times: np.ndarray) -> TransformSamples:
        """All attributes at every element of ``layer_times`` (same floats as per-time calls)."""
        layer_times = np.asarray(layer_times, dtype=np.float64)
        return TransformSamples(
            anchor_point=np.ascontiguousarray(self.anchor_point.get_values(layer_times)),
            position=np.ascontiguousarray(self.position.get_values(layer_times)),
            scale=np.ascontiguousarray(self.scale.get_values(layer_times)),
            rotation=np.ascontiguousarray(self.rotation.get_values(layer_times)[:, 0]),
            opacity=np.ascontiguousarray(self.opacity.get_values(layer_times)[:, 0]),
            origin_point=self.origin_point, blending_mode=self.blending_mode)

    def __repr__(self) -> str:
        return (f"Transform(ap={self.anchor_point}, pos={self.position}, s={self.scale}, "
                f"rot={self.rotation}, op={self.opacity}, blend={self.blending_mode})")


def _as_floats(x, n: int) -> tuple[float, ...]:
    if isinstance(x, float) or (isinstance(x, np.ndarray) and x.shape == ()):
        return (float(x),) * n
    if len(x) == 1:
        return (float(x[0]),) * n
    if len(x) == n:
        return tuple(float(v) for v in x)
    raise ValueError(f"Invalid value: {x}")


def transform_to_1dscalar(x: float | Sequence[float] | np.ndarray) -> float:
    if isinstance(x, float):
        return x
    if isinstance(x, np.ndarray) and x.shape == ():
        return float(x)
    if len(x) > 0:
        return float(x[0])
    raise ValueError(f"Invalid value: {x}")


def transform_to_2dvector(x) -> tuple[float, float]:
    return _as_floats(x, 2)  # type: ignore[return-value]


def transform_to_3dvector(x) -> tuple[float, float, float]:
    return _as_floats(x, 3)  # type: ignore[return-value]

"""Enumerations of the movis API surface that parameterise the render path.

Same names, values and ``from_string`` behaviour as the reference's ``movis/enum.py`` (BlendingMode
values double as the ``MVB_BLEND_*`` codes of the C ABI; MatteMode as ``MVB_MATTE_*``).
"""
from __future__ import annotations

import enum as _enum


class _Named(_enum.Enum):
    """Enum whose members can be looked up by their lower-case name."""

    @classmethod
    def _lookup(cls, s: str):
        if isinstance(s, str) and s == s.lower():
            member = cls.__members__.get(s.upper())
            if member is not None:
                return member
        raise ValueError(f"Unknown {cls._label()}: {s}")

    @classmethod
    def _label(cls) -> str:
        return "value"


class CacheType(_enum.Enum):
    """Kinds of render-cache entries (reference: enum.py:4-7)."""
    COMPOSITION = 0
    LAYER = 1


class AttributeType(_Named):
    """Dimension/semantics of an ``Attribute`` value (reference: enum.py:10-31)."""
    SCALAR = 0
    VECTOR2D = 1
    VECTOR3D = 2
    ANGLE = 3
    COLOR = 4

    @classmethod
    def _label(cls) -> str:
        return "attribute type"

    @staticmethod
    def from_string(s: str) -> "AttributeType":
        if s == "color":  # the reference's parser does not accept 'color' (enum.py:20-31)
            raise ValueError(f"Unknown attribute type: {s}")
        return AttributeType._lookup(s)


def _easing_members() -> dict[str, int]:
    members = {"LINEAR": 0, "EASE_IN": 1, "EASE_OUT": 2, "EASE_IN_OUT": 3, "FLAT": 4}
    orders = list(range(2, 11)) + [12, 14, 16, 18, 20, 25, 30, 35]
    for base, prefix in ((100, "EASE_IN"), (200, "EASE_OUT"), (300, "EASE_IN_OUT")):
        for n in orders:
            members[f"{prefix}{n}"] = base + n
    return members


Easing = _enum.Enum("Easing", _easing_members(), module=__name__, qualname="Easing")
Easing.__doc__ = """Interpolation curves between keyframes (reference: enum.py:34-170): LINEAR, FLAT,
EASE_IN / EASE_OUT / EASE_IN_OUT (= order 2) and EASE_*{n} for n in 2..10, 12..20 step 2, 25, 30, 35."""


def _easing_from_string(s: str) -> "Easing":
    if isinstance(s, str) and s == s.lower():
        member = Easing.__members__.get(s.upper())
        if member is not None:
            return member
    raise ValueError(f"Unknown easing: {s}")


Easing.from_string = staticmethod(_easing_from_string)  # type: ignore[attr-defined]


class BlendingMode(_Named):
    """Blend modes of ``alpha_composite`` (reference: enum.py:173-200)."""
    NORMAL = 0
    MULTIPLY = 1
    SCREEN = 2
    OVERLAY = 3
    DARKEN = 4
    LIGHTEN = 5
    COLOR_DODGE = 6
    COLOR_BURN = 7
    LINEAR_DODGE = 8
    LINEAR_BURN = 9
    HARD_LIGHT = 10
    SOFT_LIGHT = 11
    VIVID_LIGHT = 12
    LINEAR_LIGHT = 13
    PIN_LIGHT = 14
    DIFFERENCE = 15
    EXCLUSION = 16
    SUBTRACT = 17

    @classmethod
    def _label(cls) -> str:
        return "blending mode"

    @staticmethod
    def from_string(s: str) -> "BlendingMode":
        return BlendingMode._lookup(s)


class MatteMode(_Named):
    """Matte handling of ``alpha_composite`` (reference: enum.py:225-243)."""
    NONE = 0
    ALPHA = 1
    LUMINANCE = 2

    @classmethod
    def _label(cls) -> str:
        return "matte mode"

    @staticmethod
    def from_string(s: str) -> "MatteMode":
        return MatteMode._lookup(s)


class Direction(_Named):
    """Origin point of a layer (reference: enum.py:246-288)."""
    BOTTOM_LEFT = 1
    BOTTOM_CENTER = 2
    BOTTOM_RIGHT = 3
    CENTER_LEFT = 4
    CENTER = 5
    CENTER_RIGHT = 6
    TOP_LEFT = 7
    TOP_CENTER = 8
    TOP_RIGHT = 9

    @classmethod
    def _label(cls) -> str:
        return "origin point"

    @staticmethod
    def from_string(s: str) -> "Direction":
        return Direction._lookup(s)

    @staticmethod
    def to_vector(d: "Direction", size: tuple[float, float]) -> tuple[float, float]:
        """Offset of the origin point from the layer's top-left corner."""
        if not isinstance(d, Direction):
            raise ValueError(f"Unknown direction: {d}")
        col = (d.value - 1) % 3          # 0 left, 1 centre, 2 right
        row = (d.value - 1) // 3         # 0 bottom, 1 centre, 2 top
        x = (0, size[0] / 2, size[0])[col]
        y = (size[1], size[1] / 2, 0)[row]
        return (x, y)


class TextAlignment(_Named):
    """Text alignment (kept for API completeness; text rasterisation is out of scope)."""
    LEFT = 0
    CENTER = 1
    RIGHT = 2

    @classmethod
    def _label(cls) -> str:
        return "text alignment"

    @staticmethod
    def from_string(s: str) -> "TextAlignment":
        return TextAlignment._lookup(s)

"""Keyframe motion engine (host side of the render path).

API-compatible with the reference's ``movis/motion.py`` (``Motion``, ``transform_to_numpy``, the
easing table) and numerically identical to it: the interpolation is the same sequence of IEEE
float64 operations, and power easings go through the C library ``pow`` exactly as Python's
``float ** int`` does in the reference (motion.py:19-40).

Addition for the B200 path: ``Motion.sample(times)`` evaluates a whole frame range at once and
returns bit-identical values to calling the motion once per time (the per-frame
Python cost of the reference control flow caps a GPU renderer at ~300 fps).
"""
from __future__ import annotations

import bisect
import math
from typing import Callable, Sequence

import numpy as np

from .enum import AttributeType, Easing

_VALUE_DIMS = {
    AttributeType.SCALAR: 1, AttributeType.ANGLE: 1, AttributeType.VECTOR2D: 2,
    AttributeType.VECTOR3D: 3, AttributeType.COLOR: 3,
}


def transform_to_numpy(value, value_type: AttributeType) -> np.ndarray:
    """Coerce a scalar / sequence / array to the float64 vector ``value_type`` calls for."""
    dim = _VALUE_DIMS.get(value_type)
    if dim is None:
        raise ValueError(f"Invalid value type: {value_type}")
    if isinstance(value, (int, float)):
        return np.full(dim, value, dtype=np.float64)
    if isinstance(value, (Sequence, np.ndarray)) and not isinstance(value, str) and len(value) == dim:
        return np.array(value, dtype=np.float64)
    raise ValueError(f"Invalid value type: {value_type}")


class _PowerEase:
    """t**n shaped easing: kind 'in', 'out' or 'in_out' of order ``n`` (motion.py:19-40)."""

    __slots__ = ("kind", "n")

    def __init__(self, kind: str, n: int):
        self.kind = kind
        self.n = n

    def __call__(self, t: float) -> float:
        n = self.n
        if self.kind == "in":
            return t ** n
        if self.kind == "out":
            return 1.0 - (1.0 - t) ** n
        if t < 0.5:
            return 0.5 * (2 * t) ** n
        return 1.0 - 0.5 * (1.0 - 2 * (t - 0.5)) ** n

    def __repr__(self) -> str:
        return f"ease_{self.kind}{self.n}"


def linear(t: float) -> float:
    return t


def flat(t: float) -> float:
    return 0.0


def _build_easing_table() -> dict:
    table: dict = {Easing.LINEAR: linear, Easing.FLAT: flat}
    for e in Easing:
        name = e.name
        if name in ("LINEAR", "FLAT"):
            continue
        kind = "in_out" if name.startswith("EASE_IN_OUT") else ("in" if name.startswith("EASE_IN") else "out")
        digits = "".join(ch for ch in name if ch.isdigit())
        table[e] = _PowerEase(kind, int(digits) if digits else 2)
    return table


EASING_TO_FUNC: dict = _build_easing_table()


def _resolve_easing(e) -> Callable[[float], float]:
    if isinstance(e, str):
        return EASING_TO_FUNC[Easing.from_string(e)]
    if isinstance(e, Easing):
        return EASING_TO_FUNC[e]
    if callable(e):
        return e
    raise ValueError(f"Invalid easing type: {type(e)}")


def _ease_many(func: Callable[[float], float], t: np.ndarray) -> np.ndarray:
    """Apply an easing to an array with results bit-identical to scalar evaluation."""
    if func is linear:
        return t
    if func is flat:
        return np.zeros_like(t)
    # numpy's vectorised power is not guaranteed to round like libm's pow; go through Python
    # floats so every element takes the scalar code path the reference takes.  The power easings are
    # written out per kind (the same float operations as _PowerEase.__call__, without a call per element).
    tl = t.tolist()
    if type(func) is _PowerEase:
        n = func.n
        if func.kind == "in":
            out = [v ** n for v in tl]
        elif func.kind == "out":
            out = [1.0 - (1.0 - v) ** n for v in tl]
        else:
            out = [0.5 * (2 * v) ** n if v < 0.5 else 1.0 - 0.5 * (1.0 - 2 * (v - 0.5)) ** n for v in tl]
        return np.array(out, dtype=np.float64)
    return np.array([func(v) for v in tl], dtype=np.float64)


class Motion:
    """Sorted keyframes with per-segment easing for one ``Attribute`` (motion.py:103-165)."""

    def __init__(self, init_value=None, value_type: AttributeType = AttributeType.SCALAR):
        self.keyframes: list[float] = []
        self.values: list[np.ndarray] = []
        self.easings: list[Callable[[float], float]] = []
        self.value_type = value_type
        self.init_value = None if init_value is None else transform_to_numpy(init_value, value_type)

    def __len__(self) -> int:
        return len(self.keyframes)

    # -- evaluation ----------------------------------------------------------------------------
    def __call__(self, prev_value: np.ndarray, layer_time: float) -> np.ndarray:
        n = len(self.keyframes)
        if n == 0:
            if self.init_value is None:
                raise ValueError("No keyframes")
            return self.init_value
        if n == 1 or layer_time < self.keyframes[0]:
            return self.values[0]
        if self.keyframes[-1] <= layer_time:
            return self.values[-1]
        hi = bisect.bisect(self.keyframes, layer_time)
        lo = hi - 1
        span = self.keyframes[hi] - self.keyframes[lo]
        t = self.easings[lo]((layer_time - self.keyframes[lo]) / span)
        a, b = self.values[lo], self.values[hi]
        return a + (b - a) * t

    def sample(self, times: np.ndarray) -> np.ndarray:
        """Values at every element of ``times`` as an ``(N, C)`` float64 array."""
        times = np.asarray(times, dtype=np.float64)
        n = len(self.keyframes)
        if n == 0:
            if self.init_value is None:
                raise ValueError("No keyframes")
            return np.broadcast_to(self.init_value, (len(times), len(self.init_value))).copy()
        first = self.values[0]
        out = np.empty((len(times), len(first)), dtype=np.float64)
        if n == 1:
            out[:] = first
            return out
        kf = np.asarray(self.keyframes, dtype=np.float64)
        if len(times) and kf[0] <= times.min() and times.max() < kf[1]:
            # every time inside the first segment (the common case: two keyframes spanning the range):
            # the same per-element operations as below without the masks
            span = self.keyframes[1] - self.keyframes[0]
            t = _ease_many(self.easings[0], (times - self.keyframes[0]) / span)
            a, b = self.values[0], self.values[1]
            return a[None, :] + (b - a)[None, :] * t[:, None]
        hi = np.searchsorted(kf, times, side="right")
        before = times < kf[0]
        after = kf[-1] <= times
        out[before] = first
        out[after] = self.values[-1]
        inside = ~(before | after)
        for seg in np.unique(hi[inside]):
            sel = inside & (hi == seg)
            lo = int(seg) - 1
            span = self.keyframes[lo + 1] - self.keyframes[lo]
            t = _ease_many(self.easings[lo], (times[sel] - self.keyframes[lo]) / span)
            a, b = self.values[lo], self.values[lo + 1]
            out[sel] = a[None, :] + (b - a)[None, :] * t[:, None]
        return out

    # -- editing -------------------------------------------------------------------------------
    def append(self, keyframe: float, value, easing=Easing.LINEAR) -> "Motion":
        """Insert one keyframe (kept sorted); duplicates raise ``ValueError``."""
        func = _resolve_easing(easing)
        pos = bisect.bisect(self.keyframes, keyframe)
        if pos > 0 and self.keyframes[pos - 1] == keyframe:
            raise ValueError(f"Keyframe {keyframe} already exists")
        self.keyframes.insert(pos, float(keyframe))
        self.values.insert(pos, transform_to_numpy(value, self.value_type))
        self.easings.insert(pos, func)
        return self

    def extend(self, keyframes: Sequence[float], values: Sequence, easings: Sequence | None = None) -> "Motion":
        """Insert several keyframes.  ``easings`` may have ``len(keyframes)`` or one fewer entries
        (the last segment then defaults to linear); ``None`` means all linear."""
        assert len(keyframes) == len(values)
        if isinstance(easings, (str, Easing)):
            raise ValueError("easings must be a list of strings or Easing enums")
        if easings is None:
            easings = [Easing.LINEAR] * len(keyframes)
        else:
            easings = list(easings)
            if len(easings) == len(keyframes) - 1:
                easings.append(Easing.LINEAR)
            assert len(keyframes) == len(easings)
        new_times = [float(k) for k in keyframes]
        taken = set(self.keyframes)
        for k in new_times:
            if k in taken:
                raise ValueError(f"Keyframe {k} already exists")
            taken.add(k)
        new_values = [transform_to_numpy(v, self.value_type) for v in values]
        new_funcs = [_resolve_easing(e) for e in easings]
        merged = sorted(
            zip(self.keyframes + new_times, range(len(self.keyframes) + len(new_times)),
                self.values + new_values, self.easings + new_funcs),
            key=lambda item: item[0])
        self.keyframes = [m[0] for m in merged]
        self.values = [m[2] for m in merged]
        self.easings = [m[3] for m in merged]
        return self

    def clear(self) -> "Motion":
        self.keyframes, self.values, self.easings = [], [], []
        return self

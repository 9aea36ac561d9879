"""Animatable attributes (host side of the render path).

API-compatible with the reference's ``movis/attribute.py``.  ``Attribute.get_values`` is
vectorised over a frame range (bit-identical to per-time evaluation) whenever the attribute has no
user functions; with user functions it falls back to the reference's per-time loop, because those
are arbitrary Python callables ``f(prev_value, time)`` (attribute.py:161-175).
"""
from __future__ import annotations

from typing import Callable, Hashable, Sequence

import numpy as np

from .enum import AttributeType
from .motion import Motion, transform_to_numpy


class Attribute:
    """An initial value, an optional keyframe ``Motion``, optional user functions and an optional
    clip range; calling it with a layer time yields a float64 vector (attribute.py:11-75)."""

    def __init__(self, init_value, value_type: AttributeType, range: tuple[float, float] | None = None,
                 motion: Motion | None = None,
                 functions: Sequence[Callable[[np.ndarray, float], np.ndarray]] | None = None) -> None:
        value = transform_to_numpy(init_value, value_type)
        if range is not None:
            value = np.clip(value, range[0], range[1])
        self._init_value: np.ndarray = value
        self._value_type = value_type
        self._range = range
        self._motion = motion
        self._functions: list = list(functions) if functions is not None else []

    # -- evaluation ----------------------------------------------------------------------------
    @property
    def is_static(self) -> bool:
        """True when the value cannot change with time (no motion, no functions)."""
        return self._motion is None and not self._functions

    def __call__(self, layer_time: float) -> np.ndarray:
        if self.is_static:
            return transform_to_numpy(self._init_value, self._value_type)
        value = self._init_value
        if self._motion is not None:
            value = transform_to_numpy(self._motion(value, layer_time), self._value_type)
        for func in self._functions:
            value = transform_to_numpy(func(value, layer_time), self._value_type)
        if self._range is not None:
            value = np.clip(value, self._range[0], self._range[1])
        return value

    def get_values(self, layer_times: np.ndarray) -> np.ndarray:
        """``(N, C)`` array of values for an array of layer times."""
        layer_times = np.asarray(layer_times)
        assert layer_times.ndim == 1
        n, c = len(layer_times), len(self._init_value)
        if self.is_static:
            return np.broadcast_to(self._init_value.reshape(1, c), (n, c))
        if self._functions:
            if n == 0:
                return np.empty((0, c), dtype=np.float64)
            return np.concatenate([self(t) for t in layer_times]).reshape(n, c)
        values = self._motion.sample(layer_times)
        if self._range is not None:
            values = np.clip(values, self._range[0], self._range[1])
        return values

    # -- motion / functions --------------------------------------------------------------------
    def enable_motion(self) -> Motion:
        if self._motion is None:
            self._motion = Motion(init_value=self._init_value, value_type=self._value_type)
        return self._motion

    def disable_motion(self) -> None:
        self._motion = None

    @property
    def motion(self) -> Motion | None:
        return self._motion

    @property
    def functions(self) -> list:
        return self._functions

    def add_function(self, function):
        if not callable(function):
            raise ValueError(f"Invalid function: {function}")
        self._functions.append(function)
        return function

    def pop_function(self, index: int):
        return self._functions.pop(index)

    def clear_functions(self) -> None:
        self._functions.clear()

    def __getitem__(self, index: int):
        return self._functions[index]

    # -- plain properties ----------------------------------------------------------------------
    @property
    def init_value(self) -> np.ndarray:
        return self._init_value

    @init_value.setter
    def init_value(self, value) -> None:
        self.set(value)

    def set(self, init_value) -> None:
        """Replace the initial value (also that of an attached motion)."""
        value = transform_to_numpy(init_value, self._value_type)
        self._init_value = value
        if self._motion is not None:
            self._motion.init_value = value

    @property
    def value_type(self) -> AttributeType:
        return self._value_type

    @property
    def range(self) -> tuple[float, float] | None:
        return self._range

    @range.setter
    def range(self, range: tuple[float, float] | None) -> None:
        self._range = range

    def __repr__(self) -> str:
        return f"{self._init_value}" if self._motion is None else f"Attribute(value_type={self._value_type})"


class AttributesMixin:
    """Gives a layer/effect an ``attributes`` dict and a cache key built from it
    (attribute.py:193-212)."""

    @property
    def attributes(self) -> dict[str, Attribute]:
        return {k: v for k, v in vars(self).items() if isinstance(v, Attribute)}

    def get_key(self, time: float) -> tuple[Hashable, ...]:
        return tuple(transform_to_hashable(a(time)) for a in self.attributes.values())


def transform_to_hashable(x) -> float | tuple[float, ...]:
    """Scalar / length-1 vector -> float; longer vectors -> tuple of floats."""
    if isinstance(x, (int, float)):
        return float(x)
    if len(x) == 1:
        return float(x[0])
    return tuple(float(v) for v in x)

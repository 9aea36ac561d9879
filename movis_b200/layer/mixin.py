"""Timeline helper shared by sequence-like layers (reference: movis/layer/mixin.py)."""
from __future__ import annotations

from typing import Sequence

import numpy as np


class TimelineMixin:
    """Piecewise timeline of [start, end) intervals; ``get_state(t)`` finds the active one."""

    def __init__(self, start_times: Sequence[float], end_times: Sequence[float]) -> None:
        assert len(start_times) == len(end_times), f"{len(start_times)} != {len(end_times)}"
        self.start_times: np.ndarray = np.asarray(start_times, dtype=float)
        self.end_times: np.ndarray = np.asarray(end_times, dtype=float)

    def get_state(self, time: float) -> int:
        """Index of the interval containing ``time`` or -1."""
        i = int(np.searchsorted(self.end_times, time, side="right"))
        if i < len(self.start_times) and self.start_times[i] <= time < self.end_times[i]:
            return i
        return -1

    def get_states(self, times: np.ndarray) -> np.ndarray:
        """``get_state`` for an array of times (the same comparisons, vectorised)."""
        times = np.asarray(times, dtype=np.float64)
        i = np.searchsorted(self.end_times, times, side="right")
        j = np.minimum(i, len(self.start_times) - 1)
        ok = (i < len(self.start_times)) & (self.start_times[j] <= times) & (times < self.end_times[j])
        return np.where(ok, i, -1)

    @property
    def duration(self):
        return self.end_times[-1] - self.start_times[0]

"""Layer protocols of the render path (reference: movis/layer/protocol.py).

The operator interface of movis is duck-typed; these ``typing.Protocol`` classes only document it.
A layer is any callable ``time -> ndarray(H, W, 4) uint8 | None``; ``duration`` and ``get_key`` are
optional and probed with ``hasattr``.

Extension used by this package (optional, also probed with ``hasattr``): a layer may offer
``render_device(time) -> torch.Tensor | None`` returning the same image as a CUDA uint8 tensor; the
compositor then skips the host round trip.  Effects may likewise offer
``apply_device(image, time) -> torch.Tensor``.
"""
from __future__ import annotations

from typing import Hashable, Protocol

import numpy as np

AUDIO_SAMPLING_RATE = 44100
AUDIO_BLOCK_SIZE = 1024


class Layer(Protocol):
    """Minimal layer: ``__call__(time)`` returns an RGBA uint8 image or ``None``."""

    def __call__(self, time: float) -> np.ndarray | None:
        raise NotImplementedError


class BasicLayer(Protocol):
    """Layer with the optional members the compositor looks for."""

    def __call__(self, time: float) -> np.ndarray | None:
        raise NotImplementedError

    @property
    def duration(self) -> float:
        """How long the layer lasts; layers without it are treated as lasting 1e6 s."""
        return 1e6

    def get_key(self, time: float) -> Hashable:
        """Hashable state at ``time``: equal keys promise equal images (enables caching)."""
        return time


class AudioLayer(Protocol):
    """Layer that can also supply audio (audio mixing itself is outside this package's scope)."""

    def __call__(self, time: float) -> np.ndarray | None:
        raise NotImplementedError

    def get_audio(self, start_time: float, end_time: float) -> np.ndarray | None:
        raise NotImplementedError

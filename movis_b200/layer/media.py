"""Still-image sources (reference: movis/layer/media.py:18-220).

Only the sources the compositor's configs need are provided -- ``Image`` and ``ImageSequence``;
video / audio decoding is outside the render path.  Both keep their
pixels as host arrays exactly like the reference (``__call__`` hands out the same ndarray every
time) and additionally keep a device-resident copy that is uploaded once and reused by every
frame (``render_device``).
"""
from __future__ import annotations

from os import PathLike
from pathlib import Path
from typing import Sequence

import numpy as np
from PIL import Image as PILImage

from ..device import to_device, upload_source  # noqa: F401
from ..util import to_rgb
from .mixin import TimelineMixin


def _as_rgba(data: np.ndarray) -> np.ndarray:
    """Gray (H, W), RGB (H, W, 3) or RGBA (H, W, 4) uint8 -> RGBA; RGBA input is returned as is."""
    assert data.dtype == np.uint8
    if data.ndim == 2:
        data = np.repeat(data[:, :, None], 3, axis=-1)
    elif data.ndim != 3:
        raise ValueError(f"Invalid img_file shape: {data.shape}")
    if data.shape[2] == 4:
        return data
    if data.shape[2] == 3:
        alpha = np.full(data.shape[:2] + (1,), 255, dtype=np.uint8)
        return np.concatenate([data, alpha], axis=-1)
    raise ValueError(f"Invalid img_file shape: {data.shape}. Must be (H, W, 3) or (H, W, 4).")


def _load_rgba(path) -> np.ndarray:
    return np.asarray(PILImage.open(str(path)).convert("RGBA"))


class Image:
    """Still image layer: file path, ``PIL.Image`` or uint8 ndarray, shown for ``duration`` s."""

    def __init__(self, img_file: str | PathLike | PILImage.Image | np.ndarray, duration: float = 1e6) -> None:
        self._image: np.ndarray | None = None
        self._img_file: Path | None = None
        self._device_image = None
        if isinstance(img_file, (str, PathLike)):
            self._img_file = Path(img_file)
            assert self._img_file.exists(), f"{self._img_file} does not exist"
        elif isinstance(img_file, PILImage.Image):
            self._image = np.asarray(img_file.convert("RGBA"))
        elif isinstance(img_file, np.ndarray):
            self._image = _as_rgba(img_file)
        else:
            raise ValueError(f"Invalid img_file type: {type(img_file)}")
        self._duration = duration

    @classmethod
    def from_color(cls, size: tuple[int, int], color: str | tuple[int, int, int], duration: float = 1e6) -> "Image":
        """Opaque single-colour image of ``size = (width, height)``."""
        assert size[0] > 0 and size[1] > 0
        pixels = np.empty((size[1], size[0], 4), dtype=np.uint8)
        pixels[:, :, :3] = np.array(to_rgb(color)).reshape(1, 1, 3)
        pixels[:, :, 3] = 255
        return cls(pixels, duration=duration)

    @property
    def image(self) -> np.ndarray | None:
        return self._read_image()

    @property
    def duration(self) -> float:
        return self._duration

    @property
    def size(self) -> tuple[int, int]:
        h, w = self._read_image().shape[:2]
        return w, h

    def get_key(self, time: float) -> bool:
        return 0 <= time < self.duration

    def _read_image(self) -> np.ndarray:
        if self._image is None:
            assert self._img_file is not None
            self._image = _load_rgba(self._img_file)
        return self._image

    def __call__(self, time: float) -> np.ndarray | None:
        return self._read_image() if 0 <= time < self.duration else None

    def render_device(self, time: float):
        """Same image as ``__call__`` as a CUDA tensor, uploaded on first use."""
        if not (0 <= time < self.duration):
            return None
        if self._device_image is None:
            self._device_image = upload_source(self._read_image())
        return self._device_image

    def __getstate__(self):  # device tensors stay behind when a scene is shipped to a worker
        state = self.__dict__.copy()
        state["_device_image"] = None
        return state


class ImageSequence(TimelineMixin):
    """Sequence of still images, each shown during its ``[start, end)`` interval."""

    @classmethod
    def from_files(cls, img_files: Sequence[str | PathLike | PILImage.Image | np.ndarray],
                   each_duration: float = 1.0) -> "ImageSequence":
        starts = np.arange(len(img_files)) * each_duration
        return cls(starts.tolist(), (starts + each_duration).tolist(), img_files)

    @classmethod
    def from_dir(cls, img_dir: str | PathLike, each_duration: float = 1.0) -> "ImageSequence":
        exts = {".png", ".jpg", ".jpeg", ".bmp", ".tiff"}
        files = [p for p in sorted(Path(img_dir).glob("*")) if p.is_file() and p.suffix in exts]
        return cls.from_files(files, each_duration)

    def __init__(self, start_times: Sequence[float], end_times: Sequence[float],
                 img_files: Sequence[str | PathLike | PILImage.Image | np.ndarray]) -> None:
        super().__init__(start_times, end_times)
        self.img_files = img_files
        self.images: list[np.ndarray | None] = [None] * len(img_files)
        self._device_images: list = [None] * len(img_files)
        for i, f in enumerate(img_files):
            if isinstance(f, (str, PathLike)):
                assert Path(f).exists(), f"{f} does not exist"
            elif isinstance(f, PILImage.Image):
                self.images[i] = np.asarray(f.convert("RGBA"))
            elif isinstance(f, np.ndarray):
                self.images[i] = f
            else:
                raise ValueError(f"Invalid img_file type: {type(f)}")

    def get_key(self, time: float) -> int:
        i = self.get_state(time)
        return i if i >= 0 else -1

    def get_keys(self, times: np.ndarray) -> np.ndarray:
        """``get_key`` for an array of times (lets the range planner fetch each distinct image once)."""
        return self.get_states(times)

    def _host_image(self, i: int) -> np.ndarray:
        if self.images[i] is None:
            f = self.img_files[i]
            assert isinstance(f, (str, PathLike))
            self.images[i] = _load_rgba(f)
        return self.images[i]

    def __call__(self, time: float) -> np.ndarray | None:
        i = self.get_state(time)
        return None if i < 0 else self._host_image(i)

    def render_device(self, time: float):
        i = self.get_state(time)
        if i < 0:
            return None
        if self._device_images[i] is None:
            self._device_images[i] = upload_source(self._host_image(i))
        return self._device_images[i]

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_device_images"] = [None] * len(self.img_files)
        return state

"""Scene graph and per-frame renderer on the GPU.

Drop-in for the reference's ``movis/layer/composition.py`` on the render path: ``Composition`` and
``LayerItem`` keep their constructor arguments, attributes, cache-key semantics, draft
``preview_level`` behaviour and the duck-typed layer / effect protocols.  What changes is where
pixels are produced: ``Composition.__call__`` evaluates keyframes on the host (numpy float64, the
same operations in the same order as the reference so the affine matrices carry the same bits),
packs one ``mvb_layer`` descriptor per visible layer and issues ONE fused kernel
(``mvb_composite_stack``) that warps and blends the whole stack.  Layer sources, effect outputs,
finished frames and the render cache are CUDA tensors; nothing on this path computes pixels on the
CPU.

Citations ``composition.py:N`` refer to the reference file.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Hashable, Iterator, NamedTuple, Sequence

import numpy as np

from .. import _native
from ..attribute import Attribute, AttributeType
from ..device import (DeviceCache, check_image_array, empty_image, image_args, require_cuda,
                      stream_ptr, to_device, to_host)
from ..enum import BlendingMode, CacheType, Direction
from ..transform import Transform, TransformSamples, TransformValue


# ---------------------------------------------------------------------------------------------
# geometry: M = T1 @ SR @ T2, bbox and the bbox-relative matrix (composition.py:855-934)
# ---------------------------------------------------------------------------------------------

class AffinePlan(NamedTuple):
    """Per-frame warp set-up of one layer over N frames."""
    matrix: np.ndarray   # (N, 2, 3) affine_matrix_fixed: layer px -> bbox px
    bbox: np.ndarray     # (N, 4) int32: offset_x, offset_y, W, H
    valid: np.ndarray    # (N,) bool: False where the reference skips the layer (empty bbox)


def _layer_matrices(anchor: np.ndarray, position: np.ndarray, scale: np.ndarray, rotation: np.ndarray,
                    origin: Direction, size: tuple[int, int]) -> np.ndarray:
    """Stacked ``T1 @ SR @ T2`` (N, 3, 3).  numpy's stacked matmul performs the same per-slice
    products as the reference's 3x3 ``@`` (tests/test_host_engine.py pins that bit for bit)."""
    cx, cy = Direction.to_vector(origin, (float(size[0]), float(size[1])))
    return _layer_matrices_xy(anchor, position, scale, rotation, cx, cy)


def _layer_matrices_xy(anchor: np.ndarray, position: np.ndarray, scale: np.ndarray, rotation: np.ndarray,
                       cx, cy) -> np.ndarray:
    """``_layer_matrices`` with the origin vector given per sample (scalars or (N,) arrays), so the
    samples of several layers can share one call."""
    n = len(rotation)
    T1 = np.zeros((n, 3, 3), dtype=np.float64)
    T1[:, 0, 0] = T1[:, 1, 1] = T1[:, 2, 2] = 1.0
    SR = T1.copy()
    T2 = T1.copy()
    T1[:, 0, 2] = position[:, 0] + anchor[:, 0]
    T1[:, 1, 2] = position[:, 1] + anchor[:, 1]
    angle = (2 * np.pi * rotation) / 360
    cos_t, sin_t = np.cos(angle), np.sin(angle)
    SR[:, 0, 0] = scale[:, 0] * cos_t
    SR[:, 0, 1] = -scale[:, 0] * sin_t
    SR[:, 1, 0] = scale[:, 1] * sin_t
    SR[:, 1, 1] = scale[:, 1] * cos_t
    T2[:, 0, 2] = -anchor[:, 0] - cx
    T2[:, 1, 2] = -anchor[:, 1] - cy
    return T1 @ SR @ T2


def plan_affine(samples: TransformSamples, size: tuple[int, int], preview_level: int = 1) -> AffinePlan:
    """Vectorised ``_get_fixed_affine_matrix`` (composition.py:873-907) for a layer of
    ``size = (w, h)`` over all the frames in ``samples``."""
    cx, cy = Direction.to_vector(samples.origin_point, (float(size[0]), float(size[1])))
    return plan_affine_xy(samples.anchor_point, samples.position, samples.scale, samples.rotation, cx, cy,
                          float(size[0]), float(size[1]), preview_level)


def plan_affine_xy(anchor: np.ndarray, position: np.ndarray, scale: np.ndarray, rotation: np.ndarray, cx, cy, w, h,
                   preview_level: int = 1) -> AffinePlan:
    """``plan_affine`` on raw sample arrays; the origin vector ``(cx, cy)`` and the layer size ``(w, h)`` may
    be scalars or (N,) arrays, so one call can plan the samples of MANY layers (the cost of planning is a
    fixed number of small numpy calls, not their size).  Every sample is computed by the same float
    operations as in a call of its own."""
    M = _layer_matrices_xy(anchor, position, scale, rotation, cx, cy)
    n = len(M)
    inv_l = 1 / preview_level
    if preview_level == 1:
        A = M[:, :2]   # P is the identity: P @ M differs from M at most in the sign of zeros, which ceil / floor ignore
    else:
        P = np.array([[inv_l, 0, 0], [0, inv_l, 0], [0, 0, 1]], dtype=np.float64)
        A = (P @ M)[:, :2]
    corners = np.zeros((n, 4, 3), dtype=np.float64)          # (0, 0, 1), (0, h, 1), (w, 0, 1), (w, h, 1) per sample
    corners[:, :, 2] = 1.0
    corners[:, 1, 1] = corners[:, 3, 1] = h
    corners[:, 2, 0] = corners[:, 3, 0] = w
    cg = corners @ A.transpose(0, 2, 1)                      # (N, 4, 2)
    lo = np.ceil(cg.min(axis=1))
    hi = np.floor(cg.max(axis=1))
    wh = (hi - lo).astype(np.int32)
    off = lo.astype(np.int64)
    # The reference skips W == 0 or H == 0 (composition.py:898).  Negative sizes (scale exactly 0
    # at a fractional position) make cv2 fall back to "dsize = source size"; that glitch is the
    # one intentional divergence: such layers are skipped too.
    valid = (wh[:, 0] > 0) & (wh[:, 1] > 0)
    Pf = np.zeros((n, 3, 3), dtype=np.float64)
    Pf[:, 0, 0] = Pf[:, 1, 1] = inv_l
    Pf[:, 2, 2] = 1.0
    Pf[:, 0, 2] = -off[:, 0]
    Pf[:, 1, 2] = -off[:, 1]
    fixed = (Pf @ M)[:, :2]
    bbox = np.concatenate([off.astype(np.int32), wh], axis=1)
    return AffinePlan(np.ascontiguousarray(fixed), bbox, valid)


def _one_sample(p: TransformValue) -> TransformSamples:
    return TransformSamples(
        anchor_point=np.array([p.anchor_point], dtype=np.float64),
        position=np.array([p.position], dtype=np.float64),
        scale=np.array([p.scale], dtype=np.float64),
        rotation=np.array([p.rotation], dtype=np.float64),
        opacity=np.array([p.opacity], dtype=np.float64),
        origin_point=p.origin_point, blending_mode=p.blending_mode)


def get_affine_matrix(layer_size: tuple[int, int], p: TransformValue) -> np.ndarray:
    """2x3 matrix taking layer coordinates to composition coordinates (composition.py:855-870)."""
    s = _one_sample(p)
    return _layer_matrices(s.anchor_point, s.position, s.scale, s.rotation, s.origin_point, layer_size)[0, :2]


def fill_descriptors(out: np.ndarray, image, plan: AffinePlan, opacity: np.ndarray,
                     blending_mode: BlendingMode) -> None:
    """Write ``mvb_layer`` records (``_native.LAYER_DTYPE``) for one layer over len(out) frames."""
    ptr, w, h, stride = image_args(image)
    out["src"] = ptr
    out["src_w"] = w
    out["src_h"] = h
    out["src_stride"] = stride
    out["M"] = plan.matrix.reshape(-1, 6)
    out["bbox_x"] = plan.bbox[:, 0]
    out["bbox_y"] = plan.bbox[:, 1]
    out["bbox_w"] = np.where(plan.valid, plan.bbox[:, 2], 0)
    out["bbox_h"] = np.where(plan.valid, plan.bbox[:, 3], 0)
    out["blend_mode"] = blending_mode.value
    # Composition always composites with MatteMode.NONE, so the enum NORMAL selects PIL's
    # arithmetic and every other mode numpy's (imgproc.py:265-273).
    flags = _native.MVB_LAYER_PIL_OVER if blending_mode == BlendingMode.NORMAL else 0
    if getattr(image, "mvb_opaque", False):   # device.upload_source: alpha is 255 throughout
        flags |= _native.MVB_LAYER_OPAQUE_SOURCE
    out["flags"] = flags
    amap = getattr(image, "mvb_alpha_map", None)   # device.attach_alpha_map: blocks of alpha 0 are skipped
    out["alpha_map"] = 0 if amap is None else amap.data_ptr()
    out["opacity"] = opacity


# ---------------------------------------------------------------------------------------------
# LayerItem
# ---------------------------------------------------------------------------------------------

class LayerItem:
    """A layer plus its transform, timing and effect chain inside a composition
    (composition.py:578-852)."""

    def __init__(self, layer, name: str = "layer", transform: Transform | None = None, offset: float = 0.0,
                 start_time: float = 0.0, end_time: float | None = None, audio_level: float = 0.0,
                 visible: bool = True, audio: bool = True):
        self.layer = layer
        self.name: str = name
        self.transform: Transform = transform if transform is not None else Transform()
        self.offset: float = offset
        self.start_time: float = start_time
        self.end_time: float = end_time if end_time is not None else getattr(layer, "duration", 1e6)
        self.audio_level: Attribute = Attribute(audio_level, AttributeType.SCALAR, range=(-1000, 1000))
        self.visible: bool = visible
        self.audio: bool = audio
        self._effects: list = []

    # -- plain accessors -----------------------------------------------------------------------
    @property
    def duration(self) -> float:
        """``end_time - start_time`` (not the wrapped layer's own duration)."""
        return self.end_time - self.start_time

    def add_effect(self, effect):
        self._effects.append(effect)
        return effect

    def remove_effect(self, effect) -> None:
        self._effects.remove(effect)

    @property
    def effects(self) -> list:
        return self._effects

    anchor_point = property(lambda self: self.transform.anchor_point)
    position = property(lambda self: self.transform.position)
    scale = property(lambda self: self.transform.scale)
    rotation = property(lambda self: self.transform.rotation)
    opacity = property(lambda self: self.transform.opacity)
    origin_point = property(lambda self: self.transform.origin_point)
    blending_mode = property(lambda self: self.transform.blending_mode)

    def in_window(self, time: float) -> bool:
        """The time-window test of composition.py:798-800 / :189-190."""
        layer_time = time - self.offset
        return not (layer_time < self.start_time or self.end_time <= layer_time)

    def get_composition_coords(self, layer_coords: np.ndarray, time: float = 0.0,
                               layer_size: tuple[int, int] | None = None) -> np.ndarray:
        """Map layer coordinates ``(N, 2)`` or ``(2,)`` to composition coordinates ``(N, 2)``."""
        layer_time = time - self.offset
        p = self.transform.get_current_value(layer_time)
        if layer_size is None:
            result = self.layer(layer_time)
            assert result is not None, "layer is empty at the given time"
            layer_size = (result.shape[1], result.shape[0])
        matrix = get_affine_matrix(layer_size, p)
        pts = np.array(layer_coords, dtype=np.float64)
        assert pts.ndim in (1, 2), "layer_coords must be 1D or 2D"
        pts = pts.reshape(1, -1) if pts.ndim == 1 else pts
        assert pts.shape[1] == 2, "layer_coords must have 2 columns"
        homogeneous = np.concatenate([pts, np.ones((len(pts), 1))], axis=1)
        return homogeneous @ matrix.transpose()

    # -- cache keys ----------------------------------------------------------------------------
    def content_key(self, time: float) -> tuple[Hashable, Hashable]:
        """(layer key, effects key): everything the post-effect pixels depend on."""
        layer_time = time - self.offset
        layer_key = self.layer.get_key(layer_time) if hasattr(self.layer, "get_key") else layer_time
        if not self._effects:
            return layer_key, None
        return layer_key, tuple(e.get_key(layer_time) if hasattr(e, "get_key") else layer_time
                                for e in self._effects)

    def get_key(self, time: float) -> tuple[Hashable, Hashable, Hashable]:
        """(transform value, layer key, effects key) -- composition.py:752-771."""
        if not self.visible:
            return (None, None, None)
        layer_key, effects_key = self.content_key(time)
        return (self.transform.get_current_value(time - self.offset), layer_key, effects_key)

    @property
    def cacheable(self) -> bool:
        """Whether equal keys can recur: the layer and every effect define ``get_key``."""
        return hasattr(self.layer, "get_key") and all(hasattr(e, "get_key") for e in self._effects)

    # -- rendering -----------------------------------------------------------------------------
    def render_device(self, time: float):
        """Source image + effect chain as a CUDA tensor (composition.py:822-831), or ``None``."""
        if not self.visible:
            return None
        layer_time = time - self.offset
        layer = self.layer
        if hasattr(layer, "render_device"):
            frame = layer.render_device(layer_time)
            if frame is None:
                return None
        else:
            host = layer(layer_time)
            if host is None:
                return None
            check_image_array(host, "Rendered layer image")
            frame = to_device(host)
        return self.apply_effects_device(frame, layer_time)

    def apply_effects_device(self, frame, layer_time: float):
        """The effect chain of composition.py:829-830 on a CUDA image."""
        for effect in self._effects:
            if hasattr(effect, "apply_device"):
                frame = effect.apply_device(frame, layer_time)
            else:
                # foreign effect: a callable on numpy arrays (effect/protocol.py:9) -- bridge it
                host = effect(to_host(frame), layer_time)
                check_image_array(host, "Effect output")
                frame = to_device(host)
        return frame

    def __call__(self, time: float) -> np.ndarray | None:
        frame = self.render_device(time)
        return None if frame is None else to_host(frame)

    def fg_image_device(self, time: float, cache: DeviceCache | None = None):
        """Post-effect image through the layer cache.  The key leaves the transform out (pixels do
        not depend on it), so an animated layer hits the cache on every frame -- the reference
        keys on the transform too and re-stores identical pixels (composition.py:773-789)."""
        if not self.visible:
            return None
        if cache is None or not self.cacheable:
            return self.render_device(time)
        key = (CacheType.LAYER, self.name, self.content_key(time))
        hit = cache.get(key)
        if hit is not None:
            return hit
        image = self.render_device(time)
        if image is not None:
            cache[key] = image
        return image

    def __repr__(self) -> str:
        return (f"LayerItem(name={self.name!r}, layer={self.layer!r}, transform={self.transform!r}, "
                f"offset={self.offset}, visible={self.visible})")


# ---------------------------------------------------------------------------------------------
# Composition
# ---------------------------------------------------------------------------------------------

class Composition:
    """Ordered stack of layers rendered into one RGBA frame per time (composition.py:26-370).

    Args:
        size: ``(width, height)`` of the composition.
        duration: length of the composition in seconds.
        cache_size_limit: byte budget of the device-resident render cache (the reference uses a
            1 GiB disk cache).
    """

    def __init__(self, size: tuple[int, int] = (1920, 1080), duration: float = 1.0,
                 cache_size_limit: int = 1024 * 1024 * 1024) -> None:
        assert duration > 0, "duration must be positive"
        assert isinstance(size, tuple) and len(size) == 2, "size must be a tuple of length 2"
        assert size[0] > 0 and size[1] > 0, "size must be positive"
        self._layers: list[LayerItem] = []
        self._name_to_layer: dict[str, LayerItem] = {}
        self._duration = duration
        self._size = size
        self._preview_level: int = 1
        self._cache = DeviceCache(cache_size_limit)

    # -- properties ----------------------------------------------------------------------------
    @property
    def size(self) -> tuple[int, int]:
        return self._size

    @property
    def duration(self) -> float:
        return self._duration

    @property
    def preview_level(self) -> int:
        """Draft divisor: the frame is rendered at ``(W // level, H // level)``."""
        return self._preview_level

    @preview_level.setter
    def preview_level(self, level: int) -> None:
        assert level > 0, "preview_level must be greater than 0"
        self._preview_level = int(level)
        if self._preview_level != level:
            self._cache.clear()

    @contextmanager
    def preview(self, level: int = 2) -> Iterator[None]:
        """``with comp.preview(level=2): ...`` renders drafts inside the block."""
        assert level > 0
        saved = self.preview_level
        self.preview_level = level
        try:
            yield
        finally:
            self.preview_level = saved

    @property
    def cache(self) -> DeviceCache:
        return self._cache

    # -- container protocol --------------------------------------------------------------------
    @property
    def layers(self) -> Sequence[LayerItem]:
        return self._layers

    def keys(self) -> list[str]:
        return [item.name for item in self._layers]

    def values(self) -> list[LayerItem]:
        return self._layers

    def items(self) -> list[tuple[str, LayerItem]]:
        return [(item.name, item) for item in self._layers]

    def __len__(self) -> int:
        return len(self._layers)

    def __getitem__(self, key: str) -> LayerItem:
        return self._name_to_layer[key]

    def __contains__(self, key: str) -> bool:
        return key in self._name_to_layer

    def __setitem__(self, key: str, value) -> None:
        if isinstance(value, LayerItem):
            self._layers.append(value)
            self._name_to_layer[key] = value
        elif callable(value):
            self.add_layer(value, name=key)
        else:
            raise ValueError("value must be LayerItem or Layer (i.e., callable)")

    def __delitem__(self, key: str) -> None:
        self.pop_layer(key)

    def __repr__(self) -> str:
        return f"Composition(size={self.size}, duration={self.duration}, layers={self._layers!r})"

    def add_layer(self, layer, name: str | None = None, position: tuple[float, float] | None = None,
                  scale=(1.0, 1.0), rotation: float = 0.0, opacity: float = 1.0,
                  blending_mode: BlendingMode | str = BlendingMode.NORMAL, anchor_point=(0.0, 0.0),
                  origin_point: Direction | str = Direction.CENTER, transform: Transform | None = None,
                  offset: float = 0.0, start_time: float = 0.0, end_time: float | None = None,
                  audio_level: float = 0.0, visible: bool = True, audio: bool = True) -> LayerItem:
        """Append ``layer`` (any callable ``time -> RGBA image | None``, including another
        ``Composition``) to the top of the stack; see the reference docstring
        (composition.py:199-287) for the meaning of every argument -- they are unchanged."""
        if name is None:
            name = f"layer_{len(self._layers)}"
        if name in self._name_to_layer:
            raise KeyError(f"Layer with name {name} already exists")
        if end_time is None:
            end_time = getattr(layer, "duration", 1e6)
        if transform is None:
            if position is None:
                position = self.size[0] / 2, self.size[1] / 2
            transform = Transform(position=position, scale=scale, rotation=rotation, opacity=opacity,
                                  anchor_point=anchor_point, origin_point=origin_point,
                                  blending_mode=blending_mode)
        item = LayerItem(layer, name=name, transform=transform, offset=offset, start_time=start_time,
                         end_time=end_time, audio_level=audio_level, visible=visible, audio=audio)
        self._layers.append(item)
        self._name_to_layer[name] = item
        self._cache.clear()
        return item

    def pop_layer(self, name: str) -> LayerItem:
        if name not in self._name_to_layer:
            raise KeyError(f"Layer with name {name} does not exist")
        item = self._name_to_layer.pop(name)
        self._layers.remove(item)
        self._cache.clear()
        return item

    def clear(self) -> None:
        self._layers.clear()
        self._name_to_layer.clear()
        self._cache.clear()

    # -- keys ----------------------------------------------------------------------------------
    def get_key(self, time: float) -> tuple[Hashable, ...] | None:
        """Tuple of per-layer keys describing the frame at ``time`` (composition.py:183-194)."""
        if time < 0.0 or self.duration <= time:
            return None
        keys: list[Hashable] = [CacheType.COMPOSITION]
        for item in self._layers:
            keys.append(item.get_key(time) if item.in_window(time) else None)
        return tuple(keys)

    # -- rendering -----------------------------------------------------------------------------
    def frame_shape(self) -> tuple[int, int]:
        L = self._preview_level
        return self.size[1] // L, self.size[0] // L

    def plan_frame(self, time: float):
        """Host half of one frame: ``(descriptors, keepalive)`` for ``mvb_composite_stack``.

        ``descriptors`` is a ``_native.LAYER_DTYPE`` array with one record per layer that is
        visible, inside its time window, non-empty and has a non-empty bbox, in paint order."""
        L = self._preview_level
        records = np.zeros(len(self._layers), dtype=_native.LAYER_DTYPE)
        keep = []
        n = 0
        for item in self._layers:
            if not item.in_window(time):
                continue
            image = item.fg_image_device(time, self._cache)
            if image is None:
                continue
            p = item.transform.get_current_value(time - item.offset)
            plan = plan_affine(_one_sample(p), (image.shape[1], image.shape[0]), L)
            if not plan.valid[0]:
                continue
            fill_descriptors(records[n:n + 1], image, plan, np.array([p.opacity]), p.blending_mode)
            keep.append(image)
            n += 1
        return records[:n], keep

    def render_device(self, time: float, bg_color: tuple[int, int, int, int] = (0, 0, 0, 0)):
        """The frame at ``time`` as a CUDA uint8 tensor ``(H // L, W // L, 4)`` (or ``None``
        outside ``[0, duration)``).  The tensor may be shared with the render cache: treat it as
        read-only."""
        if time < 0.0 or self.duration <= time:
            return None
        require_cuda()
        H, W = self.frame_shape()
        bg = np.asarray(bg_color, dtype=np.uint8).reshape(4)
        # bg_color is part of our key; the reference's key omits it and would hand back a frame
        # rendered with a different background (composition.py:355-361).
        key = (self.get_key(time), (H, W), bg.tobytes())
        cached = self._cache.get(key)
        if cached is not None:
            return cached
        records, keep = self.plan_frame(time)
        frame = empty_image(H, W)
        ptr, w, h, stride = image_args(frame)
        _native.check(_native.lib().mvb_composite_stack(
            ptr, w, h, stride, bg.ctypes.data, len(records), records.ctypes.data, 0, stream_ptr()))
        del keep  # sources were only needed until the launch was enqueued on this stream
        self._cache[key] = frame
        return frame

    def __call__(self, time: float, bg_color: tuple[int, int, int, int] = (0, 0, 0, 0)) -> np.ndarray | None:
        """Render the frame at ``time``; returns ``ndarray (H // L, W // L, 4) uint8`` or ``None``
        (composition.py:345-370).  The array is a fresh host copy of the device frame."""
        frame = self.render_device(time, bg_color)
        return None if frame is None else to_host(frame)

    # -- frame ranges (the caller of the hot path, composition.py:405-413) -----------------------
    def render_frames(self, times, bg_color: tuple[int, int, int, int] = (0, 0, 0, 255), out=None):
        """Render ``len(times)`` frames into one CUDA tensor ``(N, H // L, W // L, 4)`` with
        keyframes evaluated in one batch; see ``movis_b200.render.render_frames``."""
        from ..render import render_frames
        return render_frames(self, times, bg_color=bg_color, out=out)

    def _write_video(self, start_time: float, end_time: float, fps: float, writer) -> None:
        """Frame loop of ``write_video`` (composition.py:405-413): renders ``np.arange(start, end,
        1/fps)`` in batches on the GPU and hands uint8 frames to ``writer.append_data`` in order."""
        from ..render import stream_frames
        times = np.arange(start_time, end_time, 1.0 / fps)
        for frame in stream_frames(self, times, bg_color=(0, 0, 0, 255), copy=False):
            writer.append_data(frame)
        writer.close()

    def write_video(self, dst_file, start_time: float = 0.0, end_time: float | None = None,
                    codec: str = "libx264", pixelformat: str = "yuv420p", input_params=None,
                    output_params=None, fps: float = 30.0, audio: bool = True, audio_codec=None) -> None:
        """Encode the composition with imageio-ffmpeg (composition.py:415-491).  Encoding / muxing
        is outside this package's scope: it needs ``imageio`` installed, and audio is not mixed."""
        try:
            import imageio
        except ImportError as e:  # pragma: no cover - imageio is not part of the hot path
            raise ImportError("write_video needs imageio + imageio-ffmpeg for encoding") from e
        if end_time is None:
            end_time = self.duration
        writer = imageio.get_writer(uri=str(dst_file), fps=fps, codec=codec, pixelformat=pixelformat,
                                    macro_block_size=None, ffmpeg_log_level="error",
                                    input_params=input_params, output_params=output_params)
        self._write_video(start_time, end_time, fps, writer)

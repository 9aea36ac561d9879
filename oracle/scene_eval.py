"""Render a movis-API scene graph with the CPU oracle.

TEST INFRASTRUCTURE (see oracle/oracle.py).  The walker is duck-typed: it accepts Composition /
LayerItem / effect objects of either the reference package or ``movis_b200`` (same public API) and
reproduces ``Composition.__call__`` (layer/composition.py:345-370) + ``LayerItem._composite``
(:791-820) + ``LayerItem.__call__`` (:822-831) using ONLY oracle pixel operators.  Keyframe
evaluation is delegated to the scene's own ``Transform.get_current_value`` (that host engine is
pinned separately against golden values in tests/test_host_engine.py).
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def _name(obj) -> str:
    return type(obj).__name__


def _enum_name(e) -> str:
    return e if isinstance(e, str) else e.name.lower()


def _is_composition(layer) -> bool:
    return hasattr(layer, "layers") and hasattr(layer, "add_layer") and hasattr(layer, "preview_level")


def apply_effect(effect, image: np.ndarray, t: float) -> np.ndarray:
    n = _name(effect)
    if n == "GaussianBlur":
        return O.effect_gaussian_blur(image, float(effect.radius(t)[0]))
    if n == "Glow":
        return O.effect_glow(image, float(effect.radius(t)[0]), float(effect.strength(t)[0]))
    if n == "DropShadow":
        return O.effect_drop_shadow(image, float(effect.radius(t)[0]), float(effect.offset(t)[0]),
                                    float(effect.angle(t)[0]), np.asarray(effect.color(t)),
                                    float(effect.opacity(t)[0]))
    if n == "FillColor":
        return O.effect_fill_color(image, np.asarray(effect.color(t)))
    if n == "HSLShift":
        return O.effect_hsl_shift(image, float(effect.hue(t)[0]), float(effect.saturation(t)[0]),
                                  float(effect.luminance(t)[0]))
    if n == "ChromaKey":
        out = np.empty_like(image)
        lo = np.ascontiguousarray(effect._lower_color, dtype=np.uint8)
        hi = np.ascontiguousarray(effect._upper_color, dtype=np.uint8)
        img = np.ascontiguousarray(image)
        O.lib().mvo_effect_chroma_key(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                      lo.ctypes.data, hi.ctypes.data, out.ctypes.data, out.strides[0])
        return out
    # foreign effect: plain callable on numpy arrays (effect/protocol.py:9)
    return effect(image, t)


def source_image(layer, layer_time: float) -> np.ndarray | None:
    """A layer's own image at its local time, through oracle operators only (never the product's
    device path): nested compositions, matte layers, crops and stripes are evaluated here; plain
    sources (Image, ImageSequence, foreign callables) are host arrays already."""
    if _is_composition(layer):
        return render(layer, layer_time)  # nested comps use their own preview level and bg 0
    n = _name(layer)
    if n == "AlphaMatte":
        return _alpha_matte(layer, layer_time)
    if n == "LuminanceMatte":
        return _luminance_matte(layer, layer_time)
    if n == "Stripe":
        return _stripe(layer, layer_time)
    if n == "_CropLayer":
        return _crop(layer, layer_time)
    return layer(layer_time)


def layer_image(item, time: float) -> np.ndarray | None:
    """LayerItem.__call__ (composition.py:822-831)."""
    layer_time = time - item.offset
    if not item.visible:
        return None
    frame = source_image(item.layer, layer_time)
    if frame is None:
        return None
    for e in item.effects:
        frame = apply_effect(e, frame, layer_time)
    return frame


def _crop(layer, t):  # ops.py:280-296
    img = source_image(layer.layer, t)
    if img is None:
        return None
    x, y, w, h = layer.rect
    return img[y:y + h, x:x + w]


def _stripe(layer, t):  # layer/texture.py:159-191
    if t < 0 or t >= layer.duration:
        return None
    return O.stripe(tuple(layer.size), float(layer.angle(t)[0]), np.asarray(layer.color1(t)), np.asarray(layer.color2(t)),
                    float(layer.opacity1(t)[0]), float(layer.opacity2(t)[0]), float(layer.total_width(t)[0]),
                    float(layer.phase(t)[0]), float(layer.ratio(t)[0]))


def _sub(layer, t):
    return source_image(layer, t)


def _alpha_matte(layer, t):  # layer/layer_ops.py:55-67
    if t < 0 or layer.duration <= t:
        return None
    mask = _sub(layer.mask, t)
    if mask is None:
        return None
    target = _sub(layer.target, t)
    if target is None:
        return mask
    return O.alpha_composite(mask, target, opacity=float(layer.opacity(t)[0]),
                             blending_mode=_enum_name(layer.blending_mode), matte_mode="alpha", use_pil=False)


def _luminance_matte(layer, t):  # layer/layer_ops.py:99-109
    if t < 0 or layer.duration <= t:
        return None
    mask = _sub(layer.mask, t)
    if mask is None:
        return None
    target = _sub(layer.target, t)
    if target is None:
        return None
    return O.alpha_composite(mask, target, matte_mode="luminance", use_pil=False)


def frame_plan(comp, time: float) -> list[O.LayerPlan]:
    plans = []
    for item in comp.layers:
        layer_time = time - item.offset
        if layer_time < item.start_time or item.end_time <= layer_time:
            continue
        img = layer_image(item, time)
        if img is None:
            continue
        p = item.transform.get_current_value(layer_time)
        plans.append(O.LayerPlan(
            image=img,
            transform=O.TransformSample(
                anchor_point=tuple(p.anchor_point), position=tuple(p.position), scale=tuple(p.scale),
                rotation=float(p.rotation), opacity=float(p.opacity),
                origin_point=_enum_name(p.origin_point), blending_mode=_enum_name(p.blending_mode))))
    return plans


def render(comp, time: float, bg_color=(0, 0, 0, 0), preview_level: int | None = None) -> np.ndarray | None:
    """Composition.__call__ (composition.py:345-370) on the CPU oracle; no caching."""
    if time < 0.0 or comp.duration <= time:
        return None
    L = comp.preview_level if preview_level is None else preview_level
    return O.composite_frame(tuple(comp.size), frame_plan(comp, time), bg_color, L)

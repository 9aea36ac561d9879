"""Synthetic scenes for BASELINE.json's configs (S1..S5).

Every builder takes the API module as its first argument and only uses the public movis API
(``mv.layer.Composition``, ``mv.layer.Image``, ``add_layer``, ``enable_motion`` ...), so the SAME
builder constructs the scene on the reference package (for golden vectors, in the build
container), on this package (the product) and for the oracle's scene walker.  ``div`` shrinks every
pixel dimension by an integer factor so parity tests can run the same scene at a size a CPU
finishes in seconds; ``div=1`` is the configuration the benchmark is quoted on.
"""
from __future__ import annotations

from typing import Any

import numpy as np

BLEND_MODE_NAMES = [
    "normal", "multiply", "screen", "overlay", "darken", "lighten", "color_dodge", "color_burn",
    "linear_dodge", "linear_burn", "hard_light", "soft_light", "vivid_light", "linear_light",
    "pin_light", "difference", "exclusion", "subtract",
]


def synthetic_rgba(seed: int, h: int, w: int, alpha: str = "random", rim: int = 0) -> np.ndarray:
    """Uniform-random RGBA; ``alpha`` in {'random', 'opaque'}; ``rim`` px of full transparency."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    if alpha == "opaque":
        img[..., 3] = 255
    if rim > 0:
        r = min(rim, h // 2, w // 2)
        if r > 0:
            img[:r, :, 3] = 0
            img[-r:, :, 3] = 0
            img[:, :r, 3] = 0
            img[:, -r:, 3] = 0
    return img


def _ramp(attr: Any, t0: float, t1: float, v0: Any, v1: Any, easing: str = "linear") -> None:
    attr.enable_motion().extend(keyframes=[t0, t1], values=[v0, v1], easings=[easing])


def _wrapped_ramp(attr: Any, duration: float, lo: float, hi: float, phase: float) -> None:
    """lo -> hi ramp over ``duration`` shifted by ``phase`` in [0, 1) and wrapped around."""
    m = attr.enable_motion()
    if phase <= 0.0:
        m.extend([0.0, duration], [lo, hi])
        return
    v_start = lo + (hi - lo) * phase
    t_wrap = duration * (1.0 - phase)
    m.extend([0.0, t_wrap, t_wrap + 1e-6, duration], [v_start, hi, lo, v_start])


# ---------------------------------------------------------------------------------------------
# S1 (config 0): 1080p, colour background + 8 rotating/scaling NORMAL layers
# ---------------------------------------------------------------------------------------------

def build_s1(mv: Any, div: int = 1, duration: float = 5.0):
    W, H, S = 1920 // div, 1080 // div, 512 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    comp.add_layer(mv.layer.Image.from_color((W, H), (32, 64, 96)), name="bg")
    for i in range(8):
        img = synthetic_rgba(1000 + i, S, S, alpha="opaque" if i % 2 == 0 else "random")
        item = comp.add_layer(mv.layer.Image(img), name=f"l{i}")
        _ramp(item.position, 0.0, duration, ((520 + 110 * i) / div, 540 / div),
              ((1400 - 110 * i) / div, 540 / div), "ease_in_out3")
        _ramp(item.scale, 0.0, duration, (1.0, 1.0), (1.4, 1.4), "linear")
        _ramp(item.rotation, 0.0, duration, 0.0, 360.0, "ease_out5")
    return comp


def s1_bytes_per_frame(div: int = 1) -> int:
    W, H, S = 1920 // div, 1080 // div, 512 // div
    return 4 * W * H + 4 * W * H + 8 * 4 * S * S


# ---------------------------------------------------------------------------------------------
# S2 (config 1, the headline): 4K, 16 x 1080p layers cycling every blend mode, animated opacity
# ---------------------------------------------------------------------------------------------

def s2_sources(div: int = 1) -> list[np.ndarray]:
    lw, lh = 1920 // div, 1080 // div
    out = []
    for i in range(16):
        out.append(synthetic_rgba(2000 + i, lh, lw, alpha="opaque" if i % 4 == 0 else "random",
                                  rim=(64 // div) if i % 5 == 0 else 0))
    return out


def build_s2(mv: Any, div: int = 1, variant: int = 0, duration: float = 5.0, sources=None):
    W, H = 3840 // div, 2160 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    if sources is None:
        sources = s2_sources(div)
    for i in range(16):
        mode = BLEND_MODE_NAMES[(i + variant) % 18]
        item = comp.add_layer(
            mv.layer.Image(sources[i]), name=f"l{i}",
            position=((1400 + 70 * i) / div, (1040 + 5 * i) / div), blending_mode=mode)
        _ramp(item.scale, 0.0, duration, (1.0, 1.0), (1.2, 1.2))
        _ramp(item.rotation, 0.0, duration, float(2 * i - 15), float(15 - 2 * i))
        _wrapped_ramp(item.opacity, duration, 0.2, 1.0, i / 16.0)
    return comp


def s2_bytes_per_frame(div: int = 1) -> int:
    return 4 * (3840 // div) * (2160 // div) + 16 * 4 * (1920 // div) * (1080 // div)


# ---------------------------------------------------------------------------------------------
# S2d: dense variant, 16 full-frame layers, translations only
# ---------------------------------------------------------------------------------------------

def build_s2d(mv: Any, div: int = 1, variant: int = 0, duration: float = 5.0):
    W, H = 3840 // div, 2160 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    for i in range(16):
        img = synthetic_rgba(2500 + i, H, W, alpha="opaque" if i % 4 == 0 else "random")
        mode = BLEND_MODE_NAMES[(i + variant) % 18]
        item = comp.add_layer(mv.layer.Image(img), name=f"d{i}", blending_mode=mode,
                              opacity=0.3 + 0.04 * i)
        if i % 2 == 1:  # fractional drift for the odd layers, integer placement for the even ones
            _ramp(item.position, 0.0, duration, (W / 2, H / 2), (W / 2 + 0.75, H / 2 - 0.5))
    return comp


def s2d_bytes_per_frame(div: int = 1) -> int:
    return 17 * 4 * (3840 // div) * (2160 // div)


# ---------------------------------------------------------------------------------------------
# Procedural sources: animated Stripe layers under transforms and blend modes
# ---------------------------------------------------------------------------------------------

def stripe_cases() -> list[dict]:
    """Constructor arguments of the stand-alone Stripe vectors (tests/golden/stripe.npz)."""
    rng = np.random.default_rng(77)
    cases = [
        dict(size=(100, 100)),                                              # the reference's defaults
        dict(size=(64, 48), angle=30.0, color1=(10, 200, 30), color2=(250, 20, 90), opacity1=0.7,
             total_width=17.3, phase=0.2, ratio=0.4),
        dict(size=(33, 7), angle=90.0, total_width=8.0, ratio=0.0),         # all colour 1
        dict(size=(5, 9), angle=-12.5, total_width=3.0, ratio=1.0, opacity2=0.25),   # all colour 2
        dict(size=(1, 1), angle=0.0, total_width=1.0, phase=-0.5),
        dict(size=(97, 61), angle=0.0, total_width=0.75, phase=1e-3, ratio=0.005),   # edge0 clamps to 0
        dict(size=(97, 61), angle=180.0, total_width=250.0, phase=-7.25, ratio=0.995),   # edge1 clamps to 1
    ]
    for _ in range(9):
        cases.append(dict(
            size=(int(rng.integers(2, 160)), int(rng.integers(2, 120))), angle=float(rng.uniform(-400, 400)),
            color1=tuple(int(v) for v in rng.integers(0, 256, 3)), color2=tuple(int(v) for v in rng.integers(0, 256, 3)),
            opacity1=float(rng.uniform(0, 1)), opacity2=float(rng.uniform(0, 1)),
            total_width=float(rng.uniform(0.5, 90)), phase=float(rng.uniform(-3, 3)), ratio=float(rng.uniform(0.02, 0.98))))
    return cases


def build_stripes(mv: Any, div: int = 1, duration: float = 4.0):
    W, H = 640 // div, 360 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    back = mv.layer.Stripe(size=(W, H), angle=20.0, color1=(20, 40, 90), color2=(200, 220, 240), total_width=48.0 / div)
    _ramp(back.phase, 0.0, duration, 0.0, 2.0)
    comp.add_layer(back, name="back")
    modes = ["normal", "screen", "soft_light", "difference"]
    for i, mode in enumerate(modes):
        st = mv.layer.Stripe(size=(200 // div, 160 // div), angle=15.0 + 40.0 * i, color1=(250 - 60 * i, 30 + 50 * i, 80),
                             color2=(10, 200 - 40 * i, 255 - 30 * i), opacity1=1.0 - 0.2 * i, opacity2=0.35 + 0.15 * i,
                             total_width=(14.0 + 9.0 * i) / div, ratio=0.3 + 0.1 * i)
        _ramp(st.angle, 0.0, duration, 15.0 + 40.0 * i, 195.0 - 30.0 * i, "ease_in_out3")
        _ramp(st.ratio, 0.0, duration, 0.3 + 0.1 * i, 0.9 - 0.15 * i)
        _ramp(st.color1, 0.0, duration, (250 - 60 * i, 30 + 50 * i, 80), (40, 250, 10 + 60 * i))
        item = comp.add_layer(st, name=f"s{i}", position=((130 + 125 * i) / div, (120 + 40 * i) / div), blending_mode=mode)
        _ramp(item.rotation, 0.0, duration, -20.0 * i, 35.0 + 10.0 * i)
        _ramp(item.scale, 0.0, duration, (1.0, 1.0), (1.3, 0.8))
        _ramp(item.opacity, 0.0, duration, 1.0, 0.5)
    return comp


# ---------------------------------------------------------------------------------------------
# S3 (config 2): nested 1080p comp with GaussianBlur / DropShadow / Glow inside a 4K parent
# ---------------------------------------------------------------------------------------------

def build_s3(mv: Any, div: int = 1, duration: float = 5.0):
    iw, ih, S = 1920 // div, 1080 // div, 400 // div
    inner = mv.layer.Composition(size=(iw, ih), duration=duration)
    inner.add_layer(mv.layer.Image.from_color((iw, ih), (200, 180, 40)), name="ibg", opacity=0.5)
    radius = 10.0 / div
    effects = [
        mv.effect.GaussianBlur(radius=radius),
        mv.effect.DropShadow(radius=radius, offset=20.0 / div, angle=45.0, opacity=0.5),
        mv.effect.Glow(radius=radius, strength=1.0),
    ]
    for i, x in enumerate((460, 960, 1460)):
        img = synthetic_rgba(3000 + i, S, S, alpha="random", rim=S // 8)
        item = inner.add_layer(mv.layer.Image(img), name=f"e{i}", position=(x / div, 540 / div))
        _ramp(item.rotation, 0.0, duration, 0.0, 90.0)
        item.add_effect(effects[i])
    W, H = 3840 // div, 2160 // div
    outer = mv.layer.Composition(size=(W, H), duration=duration)
    outer.add_layer(mv.layer.Image.from_color((W, H), (16, 24, 32)), name="bg")
    item = outer.add_layer(inner, name="nested")
    _ramp(item.scale, 0.0, duration, (1.5, 1.5), (1.8, 1.8))
    _ramp(item.rotation, 0.0, duration, 0.0, 30.0)
    return outer


# ---------------------------------------------------------------------------------------------
# S4 (config 3): chroma-keyed frame sequence over an animated background, draft levels
# ---------------------------------------------------------------------------------------------

def s4_sequence(div: int = 1, n: int = 8) -> list[np.ndarray]:
    W, H = 3840 // div, 2160 // div
    frames = []
    for j in range(n):
        f = synthetic_rgba(4100 + j, H, W, alpha="opaque")
        f[: H // 2, :, 0] = 0
        f[: H // 2, :, 1] = 255
        f[: H // 2, :, 2] = 0
        frames.append(f)
    return frames


class LoopedLayer:
    """``layer`` repeated end to end for ``duration`` seconds: a foreign layer in the sense of the plugin
    protocol (``duration``, ``get_key``, ``__call__``).  The frame sequence of S4 lasts 8 frames; looped, the
    keyed foreground is on screen for the whole timeline ("8 pre-generated 4K frames served
    at 30 fps").  ``render_device`` is forwarded when the wrapped layer has one, so device-resident
    sources stay on the device."""

    def __init__(self, layer: Any, duration: float):
        self.layer = layer
        self.duration = duration
        if hasattr(layer, "render_device"):
            self.render_device = lambda time: (
                layer.render_device(self._local(time)) if 0 <= time < self.duration else None)

    def _local(self, time: float) -> float:
        return time % self.layer.duration

    def get_key(self, time: float):
        return self.layer.get_key(self._local(time)) if 0 <= time < self.duration else None

    def get_keys(self, times):
        """Optional batch form of ``get_key`` (this package's range planner uses it when a layer has one)."""
        times = np.asarray(times, dtype=np.float64)
        if not hasattr(self.layer, "get_keys"):
            return np.array([self.get_key(float(t)) for t in times], dtype=object)
        keys = np.asarray(self.layer.get_keys(times % self.layer.duration), dtype=object)
        keys[(times < 0) | (times >= self.duration)] = None
        return keys

    def __call__(self, time: float):
        return self.layer(self._local(time)) if 0 <= time < self.duration else None


def build_s4(mv: Any, chroma_key_cls: Any, div: int = 1, duration: float = 5.0, n_seq: int = 8):
    W, H = 3840 // div, 2160 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    bg = synthetic_rgba(4000, 2400 // div, 4200 // div, alpha="opaque")
    item = comp.add_layer(mv.layer.Image(bg), name="bg")
    _ramp(item.position, 0.0, duration, (1900 / div, 1080 / div), (1940 / div, 1080 / div))
    seq = mv.layer.ImageSequence.from_files(s4_sequence(div, n_seq), each_duration=1.0 / 30.0)
    fg = comp.add_layer(LoopedLayer(seq, duration), name="fg")
    fg.add_effect(chroma_key_cls())
    return comp


# ---------------------------------------------------------------------------------------------
# S5 (config 4): 60 s 4K timeline, 24 layers + effects, the unit of the multi-GPU frame sharding
# ---------------------------------------------------------------------------------------------

def build_s5(mv: Any, chroma_key_cls: Any, div: int = 1, duration: float = 60.0, n_seq: int = 8):
    """S2's 16 mixed-mode layers (keyframes stretched over the timeline) + S3's nested inner
    composition (blur / drop shadow / glow inside) + S4's chroma-keyed frame sequence + 6 more
    NORMAL 1080p layers: 24 layers, 1800 frames at 30 fps."""
    W, H = 3840 // div, 2160 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    sources = s2_sources(div)
    for i in range(16):
        item = comp.add_layer(
            mv.layer.Image(sources[i]), name=f"l{i}",
            position=((1400 + 70 * i) / div, (1040 + 5 * i) / div), blending_mode=BLEND_MODE_NAMES[i % 18])
        _ramp(item.scale, 0.0, duration, (1.0, 1.0), (1.2, 1.2))
        _ramp(item.rotation, 0.0, duration, float(2 * i - 15), float(15 - 2 * i))
        _wrapped_ramp(item.opacity, duration, 0.2, 1.0, i / 16.0)
    # nested inner composition of S3 (its own 1080p frame, transparent background)
    inner = build_s3(mv, div=div, duration=duration)["nested"].layer
    item = comp.add_layer(inner, name="nested", position=(W * 0.3, H * 0.35), opacity=0.9)
    _ramp(item.scale, 0.0, duration, (0.8, 0.8), (1.1, 1.1))
    _ramp(item.rotation, 0.0, duration, 0.0, 30.0)
    # chroma-keyed sequence of S4, looping every n_seq frames
    seq = mv.layer.ImageSequence.from_files(s4_sequence(div, n_seq), each_duration=duration / n_seq)
    fg = comp.add_layer(seq, name="keyed", end_time=duration, opacity=0.8, scale=(0.6, 0.6),
                        position=(W * 0.7, H * 0.68))
    fg.add_effect(chroma_key_cls())
    for i in range(6):
        img = synthetic_rgba(5000 + i, 1080 // div, 1920 // div, alpha="random", rim=(32 // div) if i % 2 else 0)
        item = comp.add_layer(mv.layer.Image(img), name=f"x{i}", opacity=0.35 + 0.1 * i,
                              position=((900 + 400 * i) / div, (500 + 230 * i) / div))
        _ramp(item.rotation, 0.0, duration, -20.0 + 8 * i, 25.0 - 6 * i)
        _ramp(item.scale, 0.0, duration, (0.7, 0.7), (1.0, 1.0), "ease_in_out2")
    return comp


# ---------------------------------------------------------------------------------------------
# generic mixed scene for parity fuzzing (overhanging layers, anisotropic scale, every mode)
# ---------------------------------------------------------------------------------------------

def build_fuzz(mv: Any, seed: int, size=(160, 90), n_layers: int = 18, duration: float = 2.0):
    rng = np.random.default_rng(seed)
    W, H = size
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    origins = ["center", "top_left", "bottom_right", "center_left", "top_center"]
    for i in range(n_layers):
        h, w = (int(v) for v in rng.integers(4, max(H, 8), 2))
        alpha = ["random", "opaque"][int(rng.integers(0, 2))]
        img = synthetic_rgba(seed * 100 + i, h, w, alpha=alpha, rim=int(rng.integers(0, 3)))
        kw = dict(
            position=(float(rng.uniform(-0.2 * W, 1.2 * W)), float(rng.uniform(-0.2 * H, 1.2 * H))),
            scale=(float(rng.uniform(0.3, 2.5)), float(rng.uniform(0.3, 2.5))),
            rotation=float(rng.uniform(-180, 180)),
            opacity=[1.0, float(rng.uniform(0, 1))][int(rng.integers(0, 2))],
            blending_mode=BLEND_MODE_NAMES[(i + seed) % 18],
            origin_point=origins[int(rng.integers(0, len(origins)))],
            anchor_point=(float(rng.uniform(-5, 5)), float(rng.uniform(-5, 5))),
        )
        if i % 6 == 5:  # integer placement, no resampling
            kw.update(position=(float(rng.integers(0, W)), float(rng.integers(0, H))), scale=(1.0, 1.0),
                      rotation=0.0, origin_point="top_left", anchor_point=(0.0, 0.0))
        item = comp.add_layer(mv.layer.Image(img), name=f"f{i}", **kw)
        if i % 3 == 0:
            _ramp(item.rotation, 0.0, duration, kw["rotation"], kw["rotation"] + 40.0, "ease_in_out2")
        if i % 4 == 1:
            _ramp(item.opacity, 0.0, duration, 0.1, 0.9, "ease_out3")
    return comp


# ---------------------------------------------------------------------------------------------
# ops.crop views and matte layers under animated transforms
# ---------------------------------------------------------------------------------------------

def build_ops(mv: Any, div: int = 1, duration: float = 2.0):
    """A cropped image (a strided view of its source), an AlphaMatte with animated opacity, a
    LuminanceMatte whose target is itself a crop, and a crop of a nested composition: each rotated /
    scaled / blended inside one parent (ops.py:280-321, layer/layer_ops.py:55-109)."""
    W, H = 320 // div, 200 // div
    comp = mv.layer.Composition(size=(W, H), duration=duration)
    comp.add_layer(mv.layer.Image(synthetic_rgba(6000, H, W, alpha="opaque")), name="bg")
    big = mv.layer.Image(synthetic_rgba(6001, 180 // div, 240 // div, alpha="random", rim=10 // div))
    cr = mv.crop(big, (30 // div, 20 // div, 150 // div, 110 // div))
    item = comp.add_layer(cr, name="crop", position=(110 / div, 80 / div), scale=(1.2, 0.9), opacity=0.8,
                          blending_mode="screen")
    _ramp(item.rotation, 0.0, duration, -25.0, 40.0, "ease_in_out2")
    mh, mw = 96 // div, 128 // div
    # Read-only, like an image decoded from a file (media.py:111): the reference's matte layers composite
    # INTO the array their mask layer returns unless it is read-only (imgproc.py:263), and a writable mask
    # would change from call to call.
    mask_px = synthetic_rgba(6002, mh, mw, alpha="random", rim=6 // div)
    mask_px.flags.writeable = False
    mask = mv.layer.Image(mask_px)
    target = mv.layer.Image(synthetic_rgba(6003, mh, mw, alpha="random"))
    am = mv.layer.AlphaMatte(mask, target, opacity=0.6, blending_mode="multiply")
    _ramp(am.opacity, 0.0, duration, 0.2, 1.0)
    item = comp.add_layer(am, name="alpha_matte", position=(220 / div, 70 / div), blending_mode="overlay")
    _ramp(item.rotation, 0.0, duration, 10.0, -50.0)
    _ramp(item.scale, 0.0, duration, (0.8, 0.8), (1.3, 1.3))
    wide = mv.layer.Image(synthetic_rgba(6004, mh + 20 // div, mw + 40 // div, alpha="random"))
    lm = mv.layer.LuminanceMatte(mask, mv.crop(wide, (17 // div, 9 // div, mw, mh)))
    item = comp.add_layer(lm, name="lum_matte", position=(90 / div, 150 / div), rotation=15.0, opacity=0.9)
    _ramp(item.position, 0.0, duration, (90 / div, 150 / div), (140 / div, 130 / div))
    inner = build_fuzz(mv, seed=5, size=(120 // div, 90 // div), n_layers=6, duration=duration)
    item = comp.add_layer(mv.crop(inner, (25 // div, 15 // div, 80 // div, 60 // div)), name="crop_comp",
                          position=(240 / div, 150 / div), blending_mode="difference")
    _ramp(item.rotation, 0.0, duration, 0.0, 90.0)
    return comp


# ---------------------------------------------------------------------------------------------
# one FULL-SIZE frame of every BASELINE.json config, pinned against the reference by digest
# (oracle/gen_golden.py writes manifest.json["full_size_digests"]; tests/test_gpu_scenes.py asserts them)
# ---------------------------------------------------------------------------------------------

def full_size_shots(mv: Any, chroma_key_cls: Any):
    """Yields (tag, build() -> composition, time); background colour (0, 0, 0, 255), draft level 1."""
    yield "s1", (lambda: build_s1(mv)), 2.5
    yield "s2v0", (lambda: build_s2(mv, variant=0)), 2.1
    yield "s2v2", (lambda: build_s2(mv, variant=2)), 3.3
    yield "s2d", (lambda: build_s2d(mv)), 1.0
    yield "s3", (lambda: build_s3(mv)), 1.7
    yield "s4", (lambda: build_s4(mv, chroma_key_cls, n_seq=2)), 0.04
    yield "s5", (lambda: build_s5(mv, chroma_key_cls, n_seq=2)), 31.0

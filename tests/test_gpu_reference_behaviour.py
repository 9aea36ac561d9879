"""The invariants the reference's own hot-path tests pin, run against this package on the GPU
(reference tree: tests/test_imgproc.py, tests/effect/*.py, tests/test_ops.py fade extremes,
tests/layer/test_composition.py, tests/layer/test_layer_ops.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import movis_b200
    return movis_b200


@pytest.mark.parametrize("position", [(0, 0), (10, 10), (-10, -10), (90, 90), (-10, 90)])
def test_alpha_composite_shape_and_dtype(mv, position):       # tests/test_imgproc.py:28-38
    rng = np.random.default_rng(0)
    for mode in mv.BlendingMode:
        bg = rng.integers(0, 256, (100, 100, 4), dtype=np.uint8)
        fg = rng.integers(0, 256, (50, 50, 4), dtype=np.uint8)
        out = mv.alpha_composite(bg, fg, position=position, blending_mode=mode)
        assert out.shape == (100, 100, 4) and out.dtype == np.uint8


def test_alpha_composite_outside(mv):                          # tests/test_imgproc.py:41-50
    rng = np.random.default_rng(1)
    for mode in mv.BlendingMode:
        bg = rng.integers(0, 256, (100, 100, 4), dtype=np.uint8)
        fg = rng.integers(0, 256, (50, 50, 4), dtype=np.uint8)
        keep = bg.copy()
        assert np.array_equal(mv.alpha_composite(bg, fg, position=(1000, 1000), blending_mode=mode), keep)


def test_alpha_matte_keeps_bg_alpha(mv):                       # tests/test_imgproc.py:53-62
    rng = np.random.default_rng(2)
    bg = rng.integers(0, 256, (100, 100, 4), dtype=np.uint8)
    fg = rng.integers(0, 256, (50, 50, 4), dtype=np.uint8)
    alpha = bg[..., 3].copy()
    out = mv.alpha_composite(bg, fg, position=(10, 10), matte_mode=mv.MatteMode.ALPHA)
    assert np.array_equal(out[..., 3], alpha)


def test_luminance_matte_black_bg(mv):                         # tests/test_imgproc.py:65-76
    bg = np.zeros((100, 100, 4), dtype=np.uint8)
    bg[..., 3] = 255
    fg = np.full((100, 100, 4), 255, dtype=np.uint8)
    out = mv.alpha_composite(bg, fg, matte_mode=mv.MatteMode.LUMINANCE)
    assert (out[..., 3] == 0).all()


def _scene(mv, image, effect=None):
    scene = mv.layer.Composition(size=(image.shape[1], image.shape[0]), duration=1.0)
    item = scene.add_layer(mv.layer.Image(image))
    if effect is not None:
        item.add_effect(effect)
    return scene


def test_identity_then_gaussian_blur(mv):                      # tests/effect/test_blur.py:5-15
    image = np.full((100, 100, 4), 255, dtype=np.uint8)
    scene = _scene(mv, image)
    assert np.array_equal(scene(0.0), image)                   # identity warp, opaque layer: exact
    scene.layers[0].add_effect(mv.effect.GaussianBlur(radius=3.0))
    scene.cache.clear()
    out = scene(0.0)
    assert (out[0, 0, :3] == 255).all() and out[0, 0, 3] < 255


def test_glow_brightens(mv):                                   # tests/effect/test_blur.py:18-27
    image = np.zeros((100, 100, 4), dtype=np.uint8)
    image[..., 3] = 255
    image[40:60, 40:60, :3] = 255
    out = _scene(mv, image, mv.effect.Glow(radius=5.0, strength=2.0))(0.0)
    assert out[50, 50, 0] > 200


def test_fill_color_and_hsl_shift(mv):                         # tests/effect/test_color.py:5-33
    image = np.zeros((10, 10, 4), dtype=np.uint8)
    image[..., 3] = 255
    assert (_scene(mv, image, mv.effect.FillColor(color=(255, 0, 0)))(0.0)[5, 5] == (255, 0, 0, 255)).all()
    image[..., 2] = 255  # blue
    assert np.array_equal(_scene(mv, image, mv.effect.HSLShift(hue=360.0))(0.0), image)
    assert (_scene(mv, image, mv.effect.HSLShift(hue=180.0))(0.0)[5, 5] == (255, 255, 0, 255)).all()


def test_drop_shadow_grows_and_keeps_centre(mv):               # tests/effect/test_style.py:16-28
    image = np.full((100, 100, 4), 255, dtype=np.uint8)
    out = mv.effect.DropShadow(radius=2.0, offset=10.0)(image, 0.0)
    assert out.shape[0] >= 100 and out.shape[1] >= 100
    assert (out[out.shape[0] // 2, out.shape[1] // 2] == 255).all()
    out0 = mv.effect.DropShadow(radius=0.0, offset=5.0, opacity=0.5)(image, 0.0)
    assert out0.shape == (100 + 2 * 4, 100 + 2 * 4, 4)


def test_opacity_extremes_through_composition(mv):             # tests/test_ops.py:137-176
    scene = mv.layer.Composition(size=(10, 10), duration=2.0)
    item = scene.add_layer(mv.layer.Image.from_color((10, 10), (10, 200, 30)))
    item.opacity.enable_motion().extend([0.0, 1.0], [0.0, 1.0])
    assert (scene(0.0) == 0).all()
    assert (scene(0.0, bg_color=(1, 2, 3, 255)) == (1, 2, 3, 255)).all()
    assert (scene(1.0)[5, 5] == (10, 200, 30, 255)).all()
    assert scene(2.0) is None and scene(-0.1) is None


def test_preview_shapes(mv):                                   # tests/layer/test_composition.py:194-240
    scene = mv.layer.Composition(size=(640, 480), duration=1.0)
    scene.add_layer(mv.layer.Image.from_color((640, 480), "white"))
    assert scene(0.0).shape == (480, 640, 4)
    for level, shape in ((2, (240, 320, 4)), (4, (120, 160, 4))):
        with scene.preview(level=level):
            assert scene(0.0).shape == shape
    assert scene(0.0).shape == (480, 640, 4)


def test_time_window_offset_visibility(mv):
    scene = mv.layer.Composition(size=(8, 8), duration=4.0)
    scene.add_layer(mv.layer.Image.from_color((8, 8), (255, 0, 0), duration=1.0), name="a", offset=1.0)
    scene.add_layer(mv.layer.Image.from_color((8, 8), (0, 255, 0)), name="b", start_time=0.5, end_time=0.75, offset=3.0)
    assert (scene(0.5) == 0).all() and (scene(1.5)[0, 0] == (255, 0, 0, 255)).all() and (scene(2.5) == 0).all()
    assert (scene(3.2) == 0).all() and (scene(3.6)[0, 0] == (0, 255, 0, 255)).all() and (scene(3.75) == 0).all()
    scene["a"].visible = False
    assert (scene(1.5) == 0).all()


def test_foreign_layers_and_effects_use_the_plugin_protocol(mv):
    """Any callable layer / effect on numpy arrays keeps working (effect/protocol.py:9)."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (40, 60, 4), dtype=np.uint8)
    calls = []

    def layer(t):
        return src if t < 1.0 else None

    def invert(prev_image, t):                    # custom effect that rewrites alpha
        calls.append(prev_image.shape)
        out = prev_image.copy()
        out[..., 3] = 255 - out[..., 3]
        return out[:, ::2]                        # changes the size, returns a non-contiguous view

    scene = mv.layer.Composition(size=(80, 50), duration=2.0)
    item = scene.add_layer(layer, rotation=20.0, scale=1.1, opacity=0.7, blending_mode="overlay")
    item.add_effect(invert)
    got = scene(0.5, bg_color=(50, 60, 70, 255))
    fg = invert(src, 0.5)
    want = O.composite_frame((80, 50), [O.LayerPlan(fg, O.TransformSample(position=(40.0, 25.0), scale=(1.1, 1.1),
                                                                            rotation=20.0, opacity=0.7,
                                                                            blending_mode="overlay"))],
                             (50, 60, 70, 255))
    assert np.array_equal(got, want) and calls[0] == (40, 60, 4)
    assert (scene(1.5, bg_color=(50, 60, 70, 255)) == (50, 60, 70, 255)).all()   # layer returned None


def test_matte_layers(mv):                                     # tests/layer/test_layer_ops.py
    from oracle import oracle as O
    rng = np.random.default_rng(4)
    m = rng.integers(0, 256, (30, 40, 4), dtype=np.uint8)
    t = rng.integers(0, 256, (30, 40, 4), dtype=np.uint8)
    mask, target = mv.layer.Image(m, duration=1.0), mv.layer.Image(t, duration=1.0)
    am = mv.layer.AlphaMatte(mask, target, opacity=0.6, blending_mode="multiply")
    assert np.array_equal(am(0.0), O.alpha_composite(m, t, opacity=0.6, blending_mode="multiply", matte_mode="alpha"))
    assert np.array_equal(am(0.0)[..., 3], m[..., 3]) and am(1.0) is None and am.duration == 1.0
    lm = mv.layer.LuminanceMatte(mask, target)
    assert np.array_equal(lm(0.0), O.alpha_composite(m, t, matte_mode="luminance", use_pil=False))
    assert am.get_key(0.0) == ((0.6,), True, True) and lm.get_key(0.5) == (True, True)
    # the matte layers as sources of the fused stack: rotated, scaled and blended inside a composition
    from oracle import scene_eval
    scene = mv.layer.Composition(size=(64, 48), duration=1.0)
    scene.add_layer(mv.layer.Image.from_color((64, 48), (90, 120, 30)), opacity=0.7)
    scene.add_layer(am, position=(30.0, 20.0), rotation=25.0, scale=(1.3, 0.8), blending_mode="screen")
    scene.add_layer(lm, position=(40.0, 30.0), rotation=-40.0, opacity=0.8)
    for bg in ((0, 0, 0, 0), (9, 9, 9, 255)):
        got, want = scene(0.0, bg_color=bg), scene_eval.render(scene, 0.0, bg)
        assert np.array_equal(got, want), np.abs(got.astype(int) - want).max()
    assert len(np.unique(scene(0.0))) > 50

"""Host side of the path (keyframe engine, enums, geometry) against reference-generated goldens,
plus the behaviours the reference's own unit tests pin (tests/test_motion.py, test_attribute.py,
test_transform.py, tests/layer/test_composition.py in the reference tree)."""
import numpy as np
import pytest

import movis_b200 as mv
from movis_b200.layer.composition import _one_sample, get_affine_matrix, plan_affine
from movis_b200.motion import transform_to_numpy
from movis_b200.transform import TransformValue, transform_to_1dscalar, transform_to_2dvector, transform_to_3dvector


def test_enums_match_reference_values():
    assert [m.value for m in mv.BlendingMode] == list(range(18))
    assert mv.BlendingMode.from_string("soft_light") is mv.BlendingMode.SOFT_LIGHT
    with pytest.raises(ValueError):
        mv.BlendingMode.from_string("no_such_mode")
    assert mv.Easing.EASE_IN_OUT35.value == 335 and mv.Easing.EASE_OUT2.value == 202 and mv.Easing.FLAT.value == 4
    assert len(mv.Easing) == 56
    assert mv.Direction.to_vector(mv.Direction.BOTTOM_RIGHT, (10.0, 4.0)) == (10.0, 4.0)
    assert mv.Direction.to_vector(mv.Direction.CENTER, (10.0, 4.0)) == (5.0, 2.0)
    assert mv.Direction.to_vector(mv.Direction.TOP_LEFT, (10.0, 4.0)) == (0, 0)
    assert mv.MatteMode.from_string("luminance") is mv.MatteMode.LUMINANCE


def test_every_easing_against_reference(golden):
    g = golden("host_engine")
    ts = g["times"]
    for name, want in zip(g["easings"], g["easing_values"]):
        a = mv.Attribute(0.0, mv.AttributeType.SCALAR)
        a.enable_motion().extend([0.0, 2.0, 5.0], [1.0, -3.5, 7.25], [str(name), str(name)])
        got = np.array([a(t)[0] for t in ts])
        assert np.array_equal(got, want), name                       # per-time, bit-exact
        assert np.array_equal(a.get_values(ts)[:, 0], want), name     # batched, bit-exact


def test_affine_plans_against_reference(golden):
    g = golden("host_engine")
    dirs = {d.value: d for d in mv.Direction}
    for row, want in zip(g["transform_inputs"], g["affine_plans"]):
        p = TransformValue((row[2], row[3]), (row[4], row[5]), (row[6], row[7]), float(row[8]), 1.0,
                           dirs[int(row[9])], mv.BlendingMode.NORMAL)
        plan = plan_affine(_one_sample(p), (int(row[0]), int(row[1])), int(row[10]))
        if want[10] == 0 or want[8] < 0 or want[9] < 0:
            assert not plan.valid[0]
        else:
            assert plan.valid[0]
            assert np.array_equal(plan.matrix[0].reshape(6), want[:6])
            assert plan.bbox[0].tolist() == [int(v) for v in want[6:10]]


def test_batched_planning_is_bit_identical_to_per_frame():
    t = mv.Transform(position=(100.5, 20.25), scale=(1.1, 0.9), rotation=33.0, opacity=0.5)
    t.position.enable_motion().extend([0.0, 5.0], [(0.0, 0.0), (500.0, 300.0)], ["ease_in_out3"])
    t.rotation.enable_motion().extend([0.0, 2.0, 5.0], [0.0, 90.0, 360.0], ["ease_out5", "ease_in7"])
    t.scale.enable_motion().extend([0.0, 5.0], [(1.0, 1.0), (1.4, 1.4)])
    t.opacity.enable_motion().extend([0.0, 5.0], [0.0, 2.0])  # clipped to [0, 1]
    times = np.arange(0, 5, 1 / 30)
    S = t.sample(times)
    P = plan_affine(S, (512, 300), 2)
    for i, tt in enumerate(times):
        p = t.get_current_value(tt)
        assert S.at(i) == p
        q = plan_affine(_one_sample(p), (512, 300), 2)
        assert np.array_equal(q.matrix[0], P.matrix[i]) and np.array_equal(q.bbox[0], P.bbox[i])
    assert S.opacity.max() == 1.0


def test_motion_api():
    m = mv.Motion(value_type=mv.AttributeType.SCALAR)
    with pytest.raises(ValueError):
        m(np.zeros(1), 0.0)
    m.append(1.0, 2.0).append(0.0, 1.0, "ease_in")
    assert m.keyframes == [0.0, 1.0] and len(m) == 2
    with pytest.raises(ValueError):
        m.append(1.0, 5.0)
    with pytest.raises(ValueError):
        m.extend([2.0, 2.0], [1.0, 2.0])
    with pytest.raises(ValueError):
        m.extend([3.0], [1.0], "linear")
    assert m(None, -1.0)[0] == 1.0 and m(None, 5.0)[0] == 2.0
    assert m(None, 0.5)[0] == 1.0 + 0.25
    m.extend([2.0, 3.0], [2.0, 3.0], ["ease_in_out"])
    assert len(m) == 4 and m.easings[3](0.3) == 0.3  # padded with linear
    f = lambda t: t * t  # noqa: E731
    m.append(4.0, 0.0, f)
    assert m.easings[-1] is f
    m.clear()
    assert len(m) == 0
    m2 = mv.Motion(init_value=(1.0, 2.0), value_type=mv.AttributeType.VECTOR2D)
    assert m2(None, 0.3).tolist() == [1.0, 2.0]


def test_attribute_api():
    a = mv.Attribute(5.0, mv.AttributeType.SCALAR, range=(0.0, 1.0))
    assert a.init_value.tolist() == [1.0]
    a.set(0.25)
    assert a(0.0).tolist() == [0.25] and a.is_static
    a.add_function(lambda v, t: v + t)
    assert a(0.5).tolist() == [0.75] and a(10.0).tolist() == [1.0]
    assert a.get_values(np.array([0.0, 0.5])).tolist() == [[0.25], [0.75]]
    a.pop_function(0)
    with pytest.raises(ValueError):
        a.add_function(3)
    mot = a.enable_motion()
    assert a.enable_motion() is mot and a.motion is mot
    a.init_value = 0.5
    assert mot.init_value.tolist() == [0.5]
    a.disable_motion()
    assert a.motion is None
    v = mv.Attribute((1, 2), mv.AttributeType.VECTOR2D)
    assert v(0).dtype == np.float64 and v.get_values(np.zeros(3)).shape == (3, 2)
    with pytest.raises(ValueError):
        mv.Attribute((1, 2, 3), mv.AttributeType.VECTOR2D)
    assert transform_to_numpy(1, mv.AttributeType.COLOR).tolist() == [1.0, 1.0, 1.0]
    assert mv.transform_to_hashable(np.array([2.0])) == 2.0
    assert mv.transform_to_hashable(np.array([2.0, 3.0])) == (2.0, 3.0)
    assert mv.transform_to_hashable(3) == 3.0

    class Fx(mv.AttributesMixin):
        def __init__(self):
            self.a = mv.Attribute(1.0, mv.AttributeType.SCALAR)
            self.b = mv.Attribute((1.0, 2.0), mv.AttributeType.VECTOR2D)
            self.c = 3
    assert list(Fx().attributes) == ["a", "b"] and Fx().get_key(0.0) == (1.0, (1.0, 2.0))


def test_transform_api():
    t = mv.Transform(position=3.0, scale=2.0, rotation=10.0, opacity=5.0, origin_point="top_left",
                     blending_mode="screen")
    p = t.get_current_value(0.0)
    assert p == TransformValue((0.0, 0.0), (3.0, 3.0), (2.0, 2.0), 10.0, 1.0, mv.Direction.TOP_LEFT,
                               mv.BlendingMode.SCREEN)
    assert hash(p) == hash(t.get_current_value(1.0))
    assert list(t.attributes) == ["anchor_point", "position", "scale", "rotation", "opacity"]
    c = mv.Transform.from_positions((640, 480))
    assert c.get_current_value(0).position == (320.0, 240.0) and c.origin_point is mv.Direction.CENTER
    assert mv.Transform.from_positions((640, 480), object_fit="cover").scale(0)[0] == 640 / 480
    assert mv.Transform.from_positions((640, 480), object_fit="contain").scale(0)[0] == 480 / 640
    br = mv.Transform.from_positions((640, 480), bottom=10, right=20)
    assert br.get_current_value(0).position == (620.0, 470.0) and br.origin_point is mv.Direction.BOTTOM_RIGHT
    tl = mv.Transform.from_positions((640, 480), top=5)
    assert tl.get_current_value(0).position == (320.0, 5.0) and tl.origin_point is mv.Direction.TOP_CENTER
    with pytest.raises(ValueError):
        mv.Transform.from_positions((640, 480), top=1, bottom=1)
    with pytest.raises(ValueError):
        mv.Transform.from_positions((640, 480), object_fit="stretch")
    assert transform_to_1dscalar(np.array([4.0])) == 4.0 and transform_to_2dvector(2.0) == (2.0, 2.0)
    assert transform_to_3dvector([1, 2, 3]) == (1.0, 2.0, 3.0)
    with pytest.raises(ValueError):
        transform_to_2dvector([1, 2, 3])


def test_composition_host_logic():
    """Scene-graph behaviour that needs no GPU (reference tests/layer/test_composition.py)."""
    comp = mv.layer.Composition(size=(640, 480), duration=5.0)
    img = mv.layer.Image.from_color((640, 480), "white")
    item = comp.add_layer(img, name="item", scale=0.5)
    assert len(comp) == 1 and "item" in comp and comp["item"] is item and comp.keys() == ["item"]
    with pytest.raises(KeyError):
        comp.add_layer(img, name="item")
    coords = item.get_composition_coords(np.array([[0, 0], [640, 480]], dtype=float), layer_size=(640, 480))
    assert np.array_equal(coords, np.array([[160.0, 120.0], [480.0, 360.0]]))
    assert get_affine_matrix((640, 480), item.transform.get_current_value(0)).shape == (2, 3)
    # keys: static scene -> identical keys over time, None outside the duration
    assert comp.get_key(0.0) == comp.get_key(4.0) and comp.get_key(5.0) is None and comp.get_key(-1) is None
    item.position.enable_motion().extend([0, 1], [(0, 0), (10, 10)])
    assert comp.get_key(0.0) != comp.get_key(0.5) and comp.get_key(1.0) == comp.get_key(2.0)
    second = comp.add_layer(mv.layer.Image.from_color((8, 8), (1, 2, 3), duration=1.0), offset=1.0)
    assert second.end_time == 1.0 and second.name == "layer_1"
    assert comp.get_key(0.5)[2] is None and comp.get_key(1.5)[2] is not None  # time window
    second.visible = False
    assert comp.get_key(1.5)[2] == (None, None, None)
    assert comp.pop_layer("layer_1") is second and len(comp) == 1
    with pytest.raises(KeyError):
        comp.pop_layer("nope")
    with comp.preview(level=4):
        assert comp.preview_level == 4 and comp.frame_shape() == (120, 160)
    assert comp.preview_level == 1
    comp["fn"] = lambda t: None
    assert comp.keys() == ["item", "fn"]
    del comp["fn"]
    comp.clear()
    assert len(comp) == 0
    with pytest.raises(AssertionError):
        mv.layer.Composition(size=(0, 10))


def test_image_sources_host_side():
    rgb = np.zeros((4, 6, 3), np.uint8)
    im = mv.layer.Image(rgb, duration=2.0)
    assert im(0.0).shape == (4, 6, 4) and im(0.0)[..., 3].min() == 255 and im(2.0) is None and im.size == (6, 4)
    assert im(0.5) is im(1.0) and im.get_key(0.1) is True and im.get_key(3.0) is False
    gray = mv.layer.Image(np.full((3, 3), 7, np.uint8))
    assert gray(0)[0, 0].tolist() == [7, 7, 7, 255]
    seq = mv.layer.ImageSequence.from_files([np.zeros((2, 2, 4), np.uint8), np.ones((2, 2, 4), np.uint8)], 0.5)
    assert seq.get_key(0.2) == 0 and seq.get_key(0.7) == 1 and seq.get_key(1.0) == -1 and seq(1.2) is None
    assert seq.duration == 1.0 and seq(0.6).max() == 1
    assert mv.to_rgb("rebeccapurple") == (102, 51, 153) and mv.to_rgb("#0a0B0c") == (10, 11, 12)
    with pytest.raises(ValueError):
        mv.to_rgb("not-a-colour")

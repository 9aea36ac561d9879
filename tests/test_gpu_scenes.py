"""Whole-frame parity of the fused CUDA path.

  * against frames rendered by the reference (tests/golden/scene_frames.npz + manifest digests) on
    scaled-down versions of BASELINE.json's configs and on fuzz scenes (every blend mode,
    overhanging layers, anisotropic scale, draft levels 1/2/4, both background alphas);
  * against the CPU oracle at BASELINE.json's FULL sizes (4K / 16 layers, nested effects,
    chroma key with draft levels), where the reference needs ~13 s per frame;
  * size-independent properties: batched == per-frame, stack chunking, cache reuse.
Bar: bit-exact (max |diff| = 0); the north-star tolerance is +-1 LSB.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from scene_zoo import golden_shots, occlusion_fuzz

pytestmark = pytest.mark.gpu


from movis_b200.scenes import BLEND_MODE_NAMES as BLEND_NAMES, synthetic_rgba as scenes_rgba  # noqa: E402


def _mv():
    import movis_b200 as mv
    from movis_b200.contrib.segmentation import ChromaKey
    return mv, ChromaKey


def test_scene_frames_match_reference_golden(golden):
    mv, ChromaKey = _mv()
    g = golden("scene_frames")
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        digests = json.load(fh)["digests"]
    for tag, comp, times, bg, levels, stored in golden_shots(mv, ChromaKey):
        for L in levels:
            for t in times:
                key = f"{tag}_L{L}_t{t:.3f}"
                with comp.preview(level=L):
                    got = comp(t, bg_color=bg)
                if stored:
                    d = np.abs(got.astype(np.int16) - g[key].astype(np.int16))
                    assert d.max() == 0, f"{key}: max |diff| {d.max()}, {int((d > 0).sum())} values differ"
                assert hashlib.sha256(got.tobytes()).hexdigest() == digests[key], key


def _assert_matches_oracle(comp, t, bg, level=1):
    from oracle import scene_eval
    with comp.preview(level=level):
        got = comp(t, bg_color=bg)
    want = scene_eval.render(comp, t, bg, preview_level=level)
    assert got.shape == want.shape
    d = np.abs(got.astype(np.int16) - want.astype(np.int16))
    assert d.max() == 0, f"t={t} L={level}: max |diff| {d.max()}, {int((d > 0).sum())} values differ"
    return got


@pytest.mark.parametrize("variant", [0, 2])
def test_s2_full_size_4k_16_layers_vs_oracle(variant):
    """BASELINE.json configs[1] at full size: 3840x2160, 16 x 1080p layers, all 18 modes over the
    two variants, animated opacity."""
    from movis_b200 import scenes
    mv, _ = _mv()
    comp = scenes.build_s2(mv, div=1, variant=variant)
    frame = _assert_matches_oracle(comp, 2.1, (0, 0, 0, 255))
    assert frame.shape == (2160, 3840, 4) and (frame[..., 3] == 255).all()
    if variant == 0:
        _assert_matches_oracle(comp, 4.9, (0, 0, 0, 0))       # transparent background: general over path
        _assert_matches_oracle(comp, 0.3, (0, 0, 0, 255), level=2)


def test_s1_full_size_vs_oracle():
    from movis_b200 import scenes
    mv, _ = _mv()
    comp = scenes.build_s1(mv, div=1)
    for t in (0.0, 2.5, 4.9):
        _assert_matches_oracle(comp, t, (0, 0, 0, 255))


def test_s3_nested_effects_full_size_vs_oracle():
    from movis_b200 import scenes
    mv, _ = _mv()
    comp = scenes.build_s3(mv, div=1)
    _assert_matches_oracle(comp, 1.7, (0, 0, 0, 255))


def test_s4_chroma_key_draft_levels_full_size_vs_oracle():
    from movis_b200 import scenes
    mv, ChromaKey = _mv()
    comp = scenes.build_s4(mv, ChromaKey, div=1, n_seq=2)
    for level in (1, 2, 4):
        f = _assert_matches_oracle(comp, 0.04, (0, 0, 0, 255), level=level)
        assert f.shape == (2160 // level, 3840 // level, 4)


def test_s2d_dense_full_size_vs_oracle():
    """17 full-frame 4K layers (integer and fractional translations): the dense variant of configs[1]."""
    from movis_b200 import scenes
    mv, _ = _mv()
    _assert_matches_oracle(scenes.build_s2d(mv, div=1), 1.0, (0, 0, 0, 255))


def test_s5_timeline_full_size_vs_oracle():
    """BASELINE.json configs[4] at full size: three frames of the 60 s / 24-layer timeline (start, middle
    with every layer visible, last frame)."""
    from movis_b200 import scenes
    mv, ChromaKey = _mv()
    comp = scenes.build_s5(mv, ChromaKey, div=1, n_seq=2)
    for t in (0.0, 31.0, 1799 / 30.0):
        _assert_matches_oracle(comp, t, (0, 0, 0, 255))


def test_full_size_frames_match_reference_digests():
    """One FULL-SIZE frame of every BASELINE.json config against the sha256 of the same frame rendered by
    the unmodified reference (oracle/gen_golden.py fullsize -> manifest.json["full_size_digests"])."""
    from movis_b200 import scenes
    mv, ChromaKey = _mv()
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        want = json.load(fh)["full_size_digests"]
    seen = 0
    for tag, build, t in scenes.full_size_shots(mv, ChromaKey):
        frame = build()(t, bg_color=(0, 0, 0, 255))
        assert list(frame.shape) == want[tag]["shape"] and t == want[tag]["time"], tag
        assert hashlib.sha256(np.ascontiguousarray(frame).tobytes()).hexdigest() == want[tag]["sha256"], tag
        seen += 1
    assert seen == len(want) == 7


def test_crop_and_matte_layers_in_a_composition_match_reference_golden(golden):
    """ops.crop (a strided device view), AlphaMatte with animated opacity, LuminanceMatte over a cropped
    target and a crop of a nested composition, each under animated transforms and blend modes: frames of
    the reference, the layers' own images, a batched range, a draft level."""
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.render import render_frames
    g = golden("ops")
    comp = scenes.build_ops(mv)
    for name in ("crop", "alpha_matte", "lum_matte", "crop_comp"):
        layer = comp[name].layer
        assert np.array_equal(layer(0.9), g["layer_" + name]), name
        dev = layer.render_device(0.9)
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), g["layer_" + name]), name
    crop_dev = comp["crop"].layer.render_device(0.9)
    assert not crop_dev.is_contiguous() and crop_dev.stride(0) == 240 * 4      # a view: no pixels moved
    assert comp["crop"].layer.get_key(0.9) == comp["crop"].layer.layer.get_key(0.9) and comp["crop"].layer.duration == 1e6
    for t in g["times"]:
        assert np.array_equal(comp(float(t)), g[f"frame_{t:.3f}"]), t
    assert np.array_equal(comp(0.9, bg_color=(5, 6, 7, 255)), g["frame_opaque_0.900"])
    batch = render_frames(comp, g["times"], bg_color=(0, 0, 0, 0))
    for i, t in enumerate(g["times"]):
        assert np.array_equal(batch[i].cpu().numpy(), g[f"frame_{t:.3f}"]), t
    comp2 = scenes.build_ops(mv, div=2)
    with comp2.preview(level=2):
        assert np.array_equal(comp2(1.3), g["frame_div2_L2"])


def _occlusion_scene(mv, size=(300, 200)):
    """Whole-pixel shifts (the one-tap path), opaque NORMAL layers at opacity 1 that hide what is below them
    (dropped from the tile lists), and the near misses of both: fractional shifts, opacity 0.999, a
    non-NORMAL mode, an opaque layer that is rotated (hides only the tiles it covers completely)."""
    from movis_b200.scenes import synthetic_rgba
    W, H = size
    comp = mv.layer.Composition(size=size, duration=2.0)
    comp.add_layer(mv.layer.Image(synthetic_rgba(8000, 150, 260, "random", rim=5)), name="under", position=(140.0, 90.0),
                   rotation=20.0, blending_mode="screen")
    comp.add_layer(mv.layer.Image(synthetic_rgba(8001, H, W, "opaque")), name="cover")                  # hides "under" everywhere
    comp.add_layer(mv.layer.Image(synthetic_rgba(8002, 120, 170, "random")), name="shifted", position=(100.0, 70.0),
                   opacity=0.7, blending_mode="overlay")                                                 # whole-pixel shift, odd size / 2 = .5 offsets
    comp.add_layer(mv.layer.Image(synthetic_rgba(8003, 120, 170, "random")), name="shifted_int", position=(185.0, 110.0),
                   origin_point="top_left", blending_mode="multiply")
    comp.add_layer(mv.layer.Image(synthetic_rgba(8004, 96, 128, "opaque")), name="patch", position=(64.0, 48.0))      # occluder over 64 x 96 .. inside the frame
    comp.add_layer(mv.layer.Image(synthetic_rgba(8005, 96, 128, "opaque")), name="patch_frac", position=(200.25, 60.5))
    comp.add_layer(mv.layer.Image(synthetic_rgba(8006, 160, 220, "opaque")), name="spin", position=(150.0, 100.0), rotation=33.0)
    comp.add_layer(mv.layer.Image(synthetic_rgba(8007, 80, 80, "opaque")), name="almost", position=(250.0, 150.0), opacity=0.999)
    comp.add_layer(mv.layer.Image(synthetic_rgba(8008, 80, 80, "opaque")), name="lighten", position=(40.0, 160.0), blending_mode="lighten")
    item = comp.add_layer(mv.layer.Image(synthetic_rgba(8009, 64, 64, "opaque")), name="mover", position=(20.0, 20.0))
    item.position.enable_motion().extend([0.0, 2.0], [(20.0, 20.0), (280.0, 180.0)])                  # whole and fractional positions, leaves the frame
    return comp


def test_whole_pixel_shifts_and_hidden_layers_vs_oracle():
    mv, _ = _mv()
    comp = _occlusion_scene(mv)
    for t in (0.0, 0.5, 1.0, 1.37, 1.99):
        for bg in ((0, 0, 0, 0), (9, 8, 7, 255)):
            _assert_matches_oracle(comp, t, bg)
    _assert_matches_oracle(comp, 1.0, (9, 8, 7, 255), level=2)
    big = _occlusion_scene(mv, size=(1000, 700))          # several whole tiles under the cover
    _assert_matches_oracle(big, 0.25, (0, 0, 0, 255))
    _assert_matches_oracle(big, 0.25, (0, 0, 0, 0))


def test_randomised_shift_and_occlusion_scenes_vs_oracle():
    """A slice of tools/fuzz_parity.py (which ran 460 seeds = 11 000 frames without a mismatch)."""
    mv, _ = _mv()
    for seed in range(300, 312):
        comp = occlusion_fuzz(mv, seed)
        for t, bg, level in ((0.0, (0, 0, 0, 0), 1), (1.3, (200, 9, 200, 255), 1), (0.6, (1, 2, 3, 255), 2)):
            if min(comp.size) >= 2 * level:
                _assert_matches_oracle(comp, t, bg, level=level)


def test_alpha_map_skips_transparent_blocks_without_changing_a_pixel():
    """mvb_alpha_map_rgba8 against numpy, and scenes of sprites with wide transparent regions (every
    blend mode, rotated, both background alphas) rendered with and without the maps and by the oracle."""
    import torch
    from movis_b200 import _native
    from movis_b200.device import attach_alpha_map, to_device
    from movis_b200.render import launch_plan, plan_range
    mv, _ = _mv()
    rng = np.random.default_rng(5)
    for h, w in ((1, 1), (31, 33), (64, 64), (100, 257)):
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        img[: h // 2, :, 3] = 0
        img[:, w // 3: w // 2, 3] = 0
        got = attach_alpha_map(to_device(img)).mvb_alpha_map.cpu().numpy()
        gh, gw = (h + 31) // 32, (w + 31) // 32
        want = np.array([[img[32 * y:32 * y + 32, 32 * x:32 * x + 32, 3].max() for x in range(gw)] for y in range(gh)], np.uint8)
        assert np.array_equal(got, want), (h, w)
    comp = mv.layer.Composition(size=(640, 400), duration=2.0)
    comp.add_layer(mv.layer.Image(scenes_rgba(7000, 400, 640, "random")), name="bg", opacity=0.9)
    for i in range(18):
        sprite = scenes_rgba(7001 + i, 300, 420, "random")
        sprite[:, :140, 3] = 0            # wide transparent margins and a transparent band
        sprite[:90, :, 3] = 0
        sprite[200:240, :, 3] = 0
        item = comp.add_layer(mv.layer.Image(sprite), name=f"s{i}", position=(120.0 + 25 * i, 90.0 + 13 * i),
                              rotation=11.0 * i, scale=0.7 + 0.05 * i, opacity=1.0 if i % 3 else 0.6,
                              blending_mode=BLEND_NAMES[i])
        item.rotation.enable_motion().extend([0.0, 2.0], [11.0 * i, 11.0 * i + 50.0])
    for t in (0.0, 1.1):
        for bg in ((0, 0, 0, 0), (30, 20, 10, 255)):
            _assert_matches_oracle(comp, t, bg)
    plan = plan_range(comp, np.array([0.3, 1.7]))
    assert (plan.descriptors["alpha_map"] != 0).sum() >= 36
    out = [torch.zeros((2, 400, 640, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for bg in ((0, 0, 0, 0), (1, 2, 3, 255)):
        launch_plan(plan, out[0], bg)
        stripped = type(plan)(plan.descriptors.copy(), plan.offsets, plan.keepalive)
        stripped.descriptors["alpha_map"] = 0
        launch_plan(stripped, out[1], bg)
        torch.cuda.synchronize()
        assert torch.equal(out[0], out[1]), bg


def test_opaque_source_flag_never_changes_a_pixel():
    """MVB_LAYER_OPAQUE_SOURCE is work elimination only: the same stack with and without the flag."""
    import torch
    from movis_b200 import _native
    from movis_b200.render import launch_plan, plan_range
    mv, _ = _mv()
    comp = _occlusion_scene(mv, size=(640, 360))
    times = np.array([0.0, 0.7, 1.5])
    plan = plan_range(comp, times)
    assert (plan.descriptors["flags"] & _native.MVB_LAYER_OPAQUE_SOURCE).any()
    out = [torch.zeros((3, 360, 640, 4), dtype=torch.uint8, device="cuda") for _ in range(2)]
    launch_plan(plan, out[0], (1, 2, 3, 255))
    plan.descriptors["flags"] &= ~_native.MVB_LAYER_OPAQUE_SOURCE
    launch_plan(plan, out[1], (1, 2, 3, 255))
    torch.cuda.synchronize()
    assert torch.equal(out[0], out[1])


def test_batched_frames_equal_per_frame_renders():
    import torch
    from movis_b200 import scenes
    mv, ChromaKey = _mv()
    for comp in (scenes.build_s2(mv, div=8), scenes.build_fuzz(mv, 3), scenes.build_s4(mv, ChromaKey, div=12),
                 scenes.build_s3(mv, div=6)):
        times = np.arange(0, min(comp.duration, 2.0), 1 / 30)[::3]
        batch = comp.render_frames(times, bg_color=(3, 2, 1, 255))
        torch.cuda.synchronize()
        for i, t in enumerate(times):
            assert np.array_equal(batch[i].cpu().numpy(), comp(float(t), bg_color=(3, 2, 1, 255))), (i, t)


def test_more_layers_than_one_launch():
    """40 layers: the stack is split across launches that continue from the stored frame."""
    from movis_b200 import scenes
    from movis_b200 import _native
    mv, _ = _mv()
    assert _native.lib().mvb_max_layers_per_launch() < 40
    comp = scenes.build_fuzz(mv, seed=11, size=(200, 120), n_layers=40)
    _assert_matches_oracle(comp, 0.4, (0, 0, 0, 0))
    _assert_matches_oracle(comp, 1.1, (7, 7, 7, 255), level=2)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 31), (64, 32), (97, 65)])
def test_tiny_and_ragged_frame_sizes_vs_oracle(size):
    """Frames smaller than one 32 x 32 tile, exactly one tile, and ragged multiples, both background
    alphas and draft levels, against the oracle."""
    from movis_b200 import scenes
    mv, _ = _mv()
    comp = scenes.build_fuzz(mv, seed=21 + size[0], size=size, n_layers=10)
    for bg in ((0, 0, 0, 0), (200, 100, 50, 255)):
        _assert_matches_oracle(comp, 0.9, bg)
    if min(size) >= 4:
        _assert_matches_oracle(comp, 0.3, (1, 2, 3, 255), level=2)


def test_stream_frames_delivers_ordered_host_frames():
    from movis_b200 import scenes
    from movis_b200.render import stream_frames
    mv, _ = _mv()
    comp = scenes.build_s1(mv, div=8)
    times = np.arange(0, 5, 1 / 30)[:21]
    got = list(stream_frames(comp, times, bg_color=(0, 0, 0, 255), batch=4))
    assert len(got) == 21
    for t, f in zip(times, got):
        assert np.array_equal(f, comp(float(t), bg_color=(0, 0, 0, 255)))

    class Writer:
        def __init__(self):
            self.frames, self.closed = [], False

        def append_data(self, f):
            self.frames.append(f.copy())

        def close(self):
            self.closed = True
    w = Writer()
    comp._write_video(0.0, 0.5, 30.0, w)   # the reference's frame loop signature (composition.py:405)
    assert w.closed and len(w.frames) == 15 and np.array_equal(w.frames[3], got[3])


def test_render_cache_is_device_resident_and_keyed_like_the_reference():
    import torch
    from movis_b200 import scenes
    mv, _ = _mv()
    comp = scenes.build_s1(mv, div=8)
    a = comp.render_device(1.0, (0, 0, 0, 255))
    assert a.is_cuda and comp.render_device(1.0, (0, 0, 0, 255)) is a          # whole-frame hit
    assert comp.render_device(1.0, (0, 0, 0, 0)) is not a                       # bg colour is part of the key
    layer_keys = [k for k in comp.cache._entries if isinstance(k, tuple) and k[0] == mv.enum.CacheType.LAYER]
    assert len(layer_keys) == 9                                                 # one entry per static source
    comp.render_device(2.0, (0, 0, 0, 255))
    assert len([k for k in comp.cache._entries if k[0] == mv.enum.CacheType.LAYER]) == 9  # animated transform: still 9
    with comp.preview(level=2):
        b = comp.render_device(1.0, (0, 0, 0, 255))
    assert tuple(b.shape) == (67, 120, 4)
    comp.add_layer(mv.layer.Image.from_color((4, 4), "red"))
    assert len(comp.cache) == 0                                                 # add_layer clears (composition.py:319)
    small = mv.layer.Composition(size=(64, 64), duration=1.0, cache_size_limit=3 * 64 * 64 * 4)
    small.add_layer(lambda t: np.full((8, 8, 4), int(t * 100), np.uint8))      # no get_key: never cached as a layer
    for i in range(6):
        small.render_device(i / 10)
    assert len(small.cache) <= 3 and small.cache.nbytes <= small.cache.size_limit


def test_stripe_layers_in_a_composition_match_reference_golden(golden):
    """Device-rendered Stripe sources (animated angle / ratio / colour) under animated transforms and
    blend modes: single frames, a batched range, and a draft level, against frames of the reference."""
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.render import render_frames
    g = golden("stripe")
    comp = scenes.build_stripes(mv)
    for t in g["times"]:
        assert np.array_equal(comp(float(t)), g[f"frame_{t:.3f}"]), t
    batch = render_frames(comp, g["times"], bg_color=(0, 0, 0, 0))
    for i, t in enumerate(g["times"]):
        assert np.array_equal(batch[i].cpu().numpy(), g[f"frame_{t:.3f}"]), t
    comp2 = scenes.build_stripes(mv, div=2)
    with comp2.preview(level=2):
        assert np.array_equal(comp2(2.2), g["frame_div2_L2"])


def test_frames_rendered_on_several_streams_are_independent():
    """mvb_composite_stack keeps one coordinate-table scratch buffer per stream: frames enqueued on
    different streams back to back (no synchronisation in between) must not disturb one another."""
    import torch
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.render import render_frames
    comp = scenes.build_s2(mv, div=4)
    times = np.linspace(0.1, 4.8, 12)
    want = render_frames(comp, times, bg_color=(0, 0, 0, 255)).cpu().numpy()
    streams = [torch.cuda.Stream() for _ in range(3)]
    H, W = comp.frame_shape()
    outs = [torch.empty((1, H, W, 4), dtype=torch.uint8, device="cuda") for _ in times]
    torch.cuda.synchronize()
    for rep in range(3):
        for i, t in enumerate(times):
            with torch.cuda.stream(streams[(i + rep) % 3]):
                render_frames(comp, times[i:i + 1], bg_color=(0, 0, 0, 255), out=outs[i])
        torch.cuda.synchronize()
        for i in range(len(times)):
            assert np.array_equal(outs[i][0].cpu().numpy(), want[i]), (rep, i)

"""The CPU oracle (oracle/mvo.c) against golden vectors produced by the reference itself.

This is what pins the oracle: every array in tests/golden/ is an output of the unmodified
reference (numpy / cv2 / PIL) recorded by oracle/gen_golden.py.  Bar: bit-exact everywhere.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import scene_eval

from conftest import GOLDEN_DIR
from scene_zoo import golden_shots


def test_blend_luts_all_18_modes(golden):
    luts = golden("blend_luts")["luts"]
    assert luts.shape == (18, 256, 256)
    for mode in range(18):
        assert np.array_equal(O.blend_lut(mode), luts[mode]), O.BLEND_MODES[mode]
    # the known answers called out in SURVEY.md 8(c)
    sl = O.blend_lut("soft_light")
    assert sl[32, 128] == 243 and sl[200, 255] == 225
    assert (O.blend_lut("vivid_light")[:, 128:] == 0).all()


def test_alpha_composite_cases(golden):
    g = golden("alpha_composite")
    meta = g["meta"]
    for i, (px, py, mode, matte) in enumerate(meta):
        out = O.alpha_composite(g[f"bg{i}"], g[f"fg{i}"], (int(px), int(py)), float(g[f"op{i}"]), int(mode), int(matte))
        assert np.array_equal(out, g[f"out{i}"]), f"case {i}: mode {mode} matte {matte}"
    out = O.alpha_composite(g["str_bg"], g["str_fg"], (3, 4), 0.6, "normal", use_pil=False)
    assert np.array_equal(out, g["str_out"])


def test_over_operators_every_alpha_pair(golden):
    g = golden("over_alpha_pairs")
    bg, fg = g["bg"], g["fg"]
    assert np.array_equal(O.alpha_composite(bg, fg, (0, 0), 1.0), g["pil_1"])
    assert np.array_equal(O.alpha_composite(bg, fg, (0, 0), 0.37), g["pil_037"])
    assert np.array_equal(O.alpha_composite(bg, fg, (0, 0), 1.0, "soft_light"), g["soft_1"])
    assert np.array_equal(O.alpha_composite(bg, fg, (0, 0), 0.37, "multiply"), g["mult_037"])


def test_warp_affine(golden):
    g = golden("warp_affine")
    for i, (M, dsize) in enumerate(zip(g["M"], g["dsize"])):
        out = O.warp_affine(g[f"src{i}"], M.reshape(2, 3), (int(dsize[0]), int(dsize[1])))
        assert np.array_equal(out, g[f"dst{i}"]), f"case {i}"


def test_gaussian_weights_known_answers():
    # Q0.8 kernels OpenCV generates
    assert O.gaussian_weights_q8(0.5, 3).tolist() == [27, 202, 27]
    assert O.gaussian_weights_q8(1.0, 5).tolist() == [14, 62, 104, 62, 14]
    w10 = O.gaussian_weights_q8(10.0, 41)
    assert w10.sum() == 256 and w10[20] == 10 and w10[:6].tolist() == [1, 2, 2, 3, 3, 3]
    w25 = O.gaussian_weights_q8(25.0, 101)
    assert w25.sum() == 256 and (w25 == 0).any()
    assert O.blur_ksize(10) == 41 and O.blur_ksize(3) == 13 and O.blur_ksize(0.1) == 3


def test_gaussian_blur(golden):
    g = golden("gaussian_blur")
    for i, r in enumerate(g["radii"]):
        k = O.blur_ksize(float(r))
        assert np.array_equal(O.gaussian_blur_u8(g[f"src4_{i}"], k, float(r)), g[f"dst4_{i}"]), f"radius {r} rgba"
        assert np.array_equal(O.gaussian_blur_u8(g[f"src1_{i}"], k, float(r)), g[f"dst1_{i}"]), f"radius {r} gray"


def test_effects(golden):
    g = golden("effects")
    for i, p in enumerate(g["params"]):
        r, st, off, ang, op = (float(v) for v in p[:5])
        col = tuple(int(v) for v in p[5:8])
        hu, sat, lum = (float(v) for v in p[8:11])
        krange = tuple(float(v) for v in p[11:14])
        img = g[f"img{i}"]
        assert np.array_equal(O.effect_gaussian_blur(img, r), g[f"blur{i}"]), f"blur {i}"
        assert np.array_equal(O.effect_glow(img, r, st), g[f"glow{i}"]), f"glow {i}"
        assert np.array_equal(O.effect_drop_shadow(img, r, off, ang, col, op), g[f"shadow{i}"]), f"shadow {i}"
        assert np.array_equal(O.effect_fill_color(img, col), g[f"fill{i}"]), f"fill {i}"
        assert np.array_equal(O.effect_hsl_shift(img, hu, sat, lum), g[f"hsl{i}"]), f"hsl {i}"
        assert np.array_equal(O.effect_chroma_key(g[f"keyed{i}"], col, krange), g[f"chroma{i}"]), f"chroma {i}"
    lo, hi = O.chroma_key_bounds()
    assert np.array_equal(np.stack([lo, hi]), g["chroma_default_bounds"])
    assert lo.tolist() == [50, 178, 178] and hi.tolist() == [70, 255, 255]


def test_colour_conversions(golden):
    g = golden("color_convert")
    assert np.array_equal(O.rgb2hsv(g["rgb"]), g["hsv"])
    assert np.array_equal(O.hsv2rgb(g["hsv_in"]), g["rgb_out"])  # includes cv2's row-tail rounding


def test_affine_geometry(golden):
    g = golden("host_engine")
    names = {1: "bottom_left", 2: "bottom_center", 3: "bottom_right", 4: "center_left", 5: "center",
             6: "center_right", 7: "top_left", 8: "top_center", 9: "top_right"}
    for row, want in zip(g["transform_inputs"], g["affine_plans"]):
        w, h = int(row[0]), int(row[1])
        p = O.TransformSample(anchor_point=(row[2], row[3]), position=(row[4], row[5]), scale=(row[6], row[7]),
                              rotation=row[8], origin_point=names[int(row[9])])
        r = O.fixed_affine(p, (w, h), int(row[10]))
        if want[10] == 0:
            assert r is None
        else:
            A, (W, H), (ox, oy) = r
            assert np.array_equal(A.reshape(6), want[:6])
            assert [ox, oy, W, H] == [int(v) for v in want[6:10]]


def test_division_shortcuts_are_exact():
    """The two division shortcuts the CUDA kernels rely on, checked exhaustively on the CPU:
    x // 255 == umulhi(x, 0x01010102) for x < 2^24, and the float-reciprocal quotient estimate is
    within +-1 of floor(n / d) for every (oa, quotient boundary) pair of the numpy 'over'."""
    x = np.arange(1 << 24, dtype=np.uint64)
    assert np.array_equal(x // 255, (x * 0x01010102) >> 32)   # div255() = umulhi(x, 0x01010102)
    d = np.arange(255, 65026, dtype=np.int64)
    rd = (np.float32(1.0) / d.astype(np.float32))
    for q in (1, 2, 127, 128, 254, 255):
        for delta in (-1, 0, 1):
            n = q * d + delta
            n = n[(n >= 0) & (n < (1 << 24))]
            dd = d[: len(n)]
            est = np.floor(n.astype(np.float32) * rd[: len(n)]).astype(np.int64)
            assert np.abs(est - n // dd).max() <= 1


def test_scene_walker_matches_reference_frames(golden):
    """Whole Composition.__call__ renders: oracle scene walker over movis_b200's scene objects
    (host keyframe engine) against frames rendered by the reference."""
    import movis_b200 as mv
    from movis_b200.contrib.segmentation import ChromaKey
    g = golden("scene_frames")
    with open(os.path.join(GOLDEN_DIR, "manifest.json")) as fh:
        digests = json.load(fh)["digests"]
    n = 0
    for tag, comp, times, bg, levels, stored in golden_shots(mv, ChromaKey):
        for L in levels:
            for t in times:
                key = f"{tag}_L{L}_t{t:.3f}"
                got = scene_eval.render(comp, t, bg, preview_level=L)
                assert hashlib.sha256(got.tobytes()).hexdigest() == digests[key], key
                if stored:
                    assert np.array_equal(got, g[key]), key
                n += 1
    assert n == 74


def test_scene_walker_crop_and_matte_layers_match_reference(golden):
    """ops.crop views and AlphaMatte / LuminanceMatte sources (ops.py:280-296, layer_ops.py:55-109)
    evaluated by the oracle's scene walker, against frames and layer images of the reference."""
    import movis_b200 as mv
    from movis_b200 import scenes
    g = golden("ops")
    comp = scenes.build_ops(mv)
    for name in ("crop", "alpha_matte", "lum_matte", "crop_comp"):
        assert np.array_equal(scene_eval.source_image(comp[name].layer, 0.9), g["layer_" + name]), name
    for t in g["times"]:
        assert np.array_equal(scene_eval.render(comp, float(t), (0, 0, 0, 0)), g[f"frame_{t:.3f}"]), t
    assert np.array_equal(scene_eval.render(comp, 0.9, (5, 6, 7, 255)), g["frame_opaque_0.900"])
    comp2 = scenes.build_ops(mv, div=2)
    assert np.array_equal(scene_eval.render(comp2, 1.3, (0, 0, 0, 0), preview_level=2), g["frame_div2_L2"])


def test_reciprocal_table_division_is_exact():
    """The fused kernel divides by d in 1..256 with umulhi(n << 8, ceil(2^24 / d)) for n <= 65025
    (color_dodge / color_burn / vivid_light, movis_b200/csrc/pixel_math.cuh): exhaustive check."""
    n = np.arange(65026, dtype=np.uint64)
    for d in range(1, 257):
        m = np.uint64(((1 << 24) + d - 1) // d)
        assert np.array_equal(((n << np.uint64(8)) * m) >> np.uint64(32), n // np.uint64(d)), d


def test_float_division_constants_of_the_stack_kernel():
    """movis_b200/csrc/stack_float.cuh divides by 255 with ONE fused multiply-add against the float32
    constant fl(1/255) = 0x3B808081 (which lies above 1/255):  fma.rm(x, c, 1.5*2^23) leaves
    floor(x*c) in the low mantissa bits, fma.rn leaves rint(x*c).  Checked here exhaustively over the
    ranges the kernel uses (the products are exact in float64, like inside an FMA):
      floor(x * c)   == x // 255             0 <= x <= 65025   (numpy over, multiply, screen)
      floor(x * 2c)  == 2x // 255            0 <= x <= 65025   (overlay, hard_light, exclusion)
      rint(x * c)    == (x + 127) // 255     0 <= x <= 65025   (PIL over on an opaque frame)
    and PIL's own rounding chain equals the closed forms the kernel evaluates."""
    c = np.float64(np.float32(1.0) / np.float32(255.0))
    assert np.float32(c).view(np.uint32) == 0x3B808081 and c > 1.0 / 255.0
    x = np.arange(65026, dtype=np.int64)
    assert np.array_equal(np.floor(x * c).astype(np.int64), x // 255)
    assert np.array_equal(np.floor(x * (2 * c)).astype(np.int64), (2 * x) // 255)
    assert np.array_equal(np.rint(x * c).astype(np.int64), (x + 127) // 255)
    frac = x * c - np.floor(x * c)
    assert np.abs(frac - 0.5).min() > 1e-3          # no ties: round-to-nearest is unambiguous
    # PIL: (((t >> 8) + t) >> 8) >> 7 with t = 128 (x + 128) is floor((x + 127) / 255) ...
    t = 128 * x + 0x4000
    assert np.array_equal((((t >> 8) + t) >> 8) >> 7, (x + 127) // 255)
    # ... and for any t below 2^24 it is (t + (t >> 8)) >> 15; the alpha chain is (u + (u >> 8)) >> 8
    t = np.arange(0, 255 * 32640 + 0x4000 + 1, dtype=np.int64)
    assert np.array_equal((((t >> 8) + t) >> 8) >> 7, (t + (t >> 8)) >> 15)
    u = np.arange(0, 65025 + 0x80 + 1, dtype=np.int64)
    assert np.array_equal(((u >> 8) + u) >> 8, (u + (u >> 8)) >> 8)
    # sa == 0 -> coef1 = 0, coef2 = 32640: the formula returns d itself (PIL's "leave the pixel alone")
    d = np.arange(256, dtype=np.int64)
    td = d * 32640 + 0x4000
    assert np.array_equal((td + (td >> 8)) >> 15, d)


def test_float_bilinear_of_the_stack_kernel_is_exact():
    """The bilinear step of stack_float.cuh: the 2 x 2 weights are 32 times the exact products
    (32 - fx)(32 - fy)... (the 32 is what masking fx in place leaves), packed as 16-bit pairs by integer
    multiplies without carries between the halves; the IDP.2A sum started at the bits of 2^23 + 32 * 512
    is the float 2^23 + 32 (512 + sum) (< 2^24), and one fma rounded down, X * 2^-15 + (1.5 * 2^23 - 256),
    leaves 1.5 * 2^23 + ((sum + 512) >> 10): cv2's result."""
    for fx in range(32):
        for fy in range(32):
            a = ((32 * fx) * 0xFFFF + 1024) & 0xFFFFFFFF
            wb = (a * fy) & 0xFFFFFFFF
            wt = ((a << 5) - wb) & 0xFFFFFFFF
            assert (wt & 0xFFFF, wt >> 16) == (32 * (32 - fx) * (32 - fy), 32 * fx * (32 - fy))
            assert (wb & 0xFFFF, wb >> 16) == (32 * (32 - fx) * fy, 32 * fx * fy)
    sums = np.arange(0, 255 * 1024 + 1, dtype=np.int64)                    # every possible weighted sum
    bits = (np.uint32(0x4B000000 + 32 * 512) + (32 * sums).astype(np.uint32))
    X = bits.view(np.float32)
    assert np.array_equal(X.astype(np.int64), (1 << 23) + 32 * (512 + sums)) and X.max() < 2 ** 24
    exact = X.astype(np.float64) / 32768.0 + (12582912.0 - 256.0)          # the fma before its one rounding
    down = np.floor(exact)                                                 # ulp is 1 in [2^23, 2^24): fma.rm = floor
    assert down.max() < 2 ** 24 and np.array_equal(down.astype(np.float32).astype(np.float64), down)
    assert np.array_equal(down.astype(np.int64) - 12582912, (sums + 512) >> 10)


def test_stripe_source_matches_reference(golden):
    """Stripe (layer/texture.py:88-191): the oracle's float64 restatement against images the reference
    produced, stand-alone and as animated layers of a composition (oracle scene walker)."""
    import movis_b200 as mv
    from movis_b200 import scenes
    g = golden("stripe")
    for i, kw in enumerate(scenes.stripe_cases()):
        kw = dict(kw)
        assert np.array_equal(O.stripe(kw.pop("size"), **kw), g[f"img{i}"]), (i, kw)
    comp = scenes.build_stripes(mv)
    for t in g["times"]:
        assert np.array_equal(scene_eval.render(comp, float(t), (0, 0, 0, 0)), g[f"frame_{t:.3f}"]), t
    comp2 = scenes.build_stripes(mv, div=2)
    assert np.array_equal(scene_eval.render(comp2, 2.2, (0, 0, 0, 0), preview_level=2), g["frame_div2_L2"])


def test_float_soft_light_of_the_stack_kernel_is_exact():
    """blend_diff2<11> (stack_float.cuh) in float32, emulated with exact float64 products: the dark branch
    divides by 65025 with one fma.rm against the float above 1/65025; the light branch keeps
    D = g_w3c(b) - b only modulo 255 * 256, split as 255 A + R, so every intermediate is an exact
    integer below 2^24.  Checked for every (b, f) against the reference's soft_light table."""
    c65025 = np.array([0x37810183], dtype=np.uint32).view(np.float32)[0]
    c255 = np.array([0x3B808081], dtype=np.uint32).view(np.float32)[0]
    assert np.float32(1.5378702300949953e-05) == c65025 and np.float32(1.0) / np.float32(255.0) == c255
    x = np.arange(256, dtype=np.uint32)
    with np.errstate(all="ignore"):   # imgproc.py:64-68, uint32 wrap-around included
        g = np.where(x < 64, ((16 * x - (12 * 255 ** 2) * x + 4 * (255 ** 2)) * x) // (255 ** 2),
                     np.sqrt(255 * x).astype(np.int32)).astype(np.uint32)
    D = ((g - x) % np.uint32(65280)).astype(np.int64)
    A, R = D // 255, D % 255
    b = np.arange(256, dtype=np.int64)[:, None]
    f = np.arange(256, dtype=np.int64)[None, :]
    N = (255 - 2 * f) * b * (255 - b)
    assert N[:, :128].max() < 2 ** 22
    dark = b - np.floor(N.astype(np.float64) * np.float64(c65025)).astype(np.int64)
    k = 2 * f - 255
    kR = k * R[:, None]
    assert kR[:, 128:].max() <= 130050 and (b + k * A[:, None] + kR // 255)[:, 128:].max() < 2 ** 22
    light = (b + k * A[:, None] + np.floor(kR.astype(np.float64) * np.float64(c255)).astype(np.int64)) % 256
    T = np.where(f < 128, dark, light)
    assert np.array_equal(T.astype(np.uint8), O.blend_lut("soft_light"))

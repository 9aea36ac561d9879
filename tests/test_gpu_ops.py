"""Parity of every C-ABI operator on the GPU: against the golden vectors recorded from the
reference, and against the CPU oracle on seeded inputs.  Bar: bit-exact (uint8/integer work)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    import torch
    from movis_b200 import _native
    from movis_b200.device import image_args, stream_ptr, to_device

    class D:
        lib = _native.lib()
        native = _native

        @staticmethod
        def up(a):
            return to_device(np.ascontiguousarray(a))

        @staticmethod
        def down(t):
            torch.cuda.synchronize()
            return t.cpu().numpy()

        args = staticmethod(image_args)
        stream = staticmethod(stream_ptr)
    assert D.lib.mvb_device_check() == 0, D.lib.mvb_last_error()
    return D


def _composite(dev, bg, fg, pos, op, mode, matte, pil=None):
    import torch
    if pil is None:
        pil = (mode == 0 and matte == 0)
    b, f = dev.up(bg), dev.up(fg)
    bp, bw, bh, bs = dev.args(b)
    fp, fw, fh, fs = dev.args(f)
    dev.native.check(dev.lib.mvb_alpha_composite(bp, bw, bh, bs, fp, fw, fh, fs, int(pos[0]), int(pos[1]),
                                                 float(op), int(mode), int(matte), 1 if pil else 0, dev.stream()))
    return dev.down(b)


def test_alpha_composite_golden(dev, golden):
    g = golden("alpha_composite")
    for i, (px, py, mode, matte) in enumerate(g["meta"]):
        out = _composite(dev, g[f"bg{i}"], g[f"fg{i}"], (px, py), float(g[f"op{i}"]), mode, matte)
        assert np.array_equal(out, g[f"out{i}"]), f"case {i}: mode {mode} matte {matte}"
    out = _composite(dev, g["str_bg"], g["str_fg"], (3, 4), 0.6, 0, 0, pil=False)
    assert np.array_equal(out, g["str_out"])


def test_over_every_alpha_pair_golden(dev, golden):
    g = golden("over_alpha_pairs")
    bg, fg = g["bg"], g["fg"]
    assert np.array_equal(_composite(dev, bg, fg, (0, 0), 1.0, 0, 0), g["pil_1"])
    assert np.array_equal(_composite(dev, bg, fg, (0, 0), 0.37, 0, 0), g["pil_037"])
    assert np.array_equal(_composite(dev, bg, fg, (0, 0), 1.0, 11, 0), g["soft_1"])
    assert np.array_equal(_composite(dev, bg, fg, (0, 0), 0.37, 1, 0), g["mult_037"])


def test_every_mode_every_channel_pair_vs_oracle(dev):
    """All 65 536 (bg, fg) channel values x 18 modes x {opaque, translucent, transparent} bg alpha."""
    from oracle import oracle as O
    b = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    f = np.repeat(np.arange(256, dtype=np.uint8)[None, :], 256, 0)
    rng = np.random.default_rng(0)
    for ba in (255, 200, 0):
        bg = np.stack([b, b, b, np.full_like(b, ba)], -1)
        fg = np.stack([f, f, f, rng.integers(0, 256, b.shape, dtype=np.uint8)], -1)
        for mode in range(18):
            for matte in (0, 1):
                want = O.alpha_composite(bg, fg, (0, 0), 0.8, mode, matte, use_pil=False)
                got = _composite(dev, bg, fg, (0, 0), 0.8, mode, matte, pil=False)
                assert np.array_equal(got, want), (ba, mode, matte)


def _stack(dev, W, H, bg_color, layers, flags=0, initial=None):
    """mvb_composite_stack with identity-warp layers: (image, (x, y), mode, pil, opacity) in paint order.
    ``initial`` + flags=1 (MVB_STACK_KEEP_FRAME) starts from given frame contents."""
    import torch
    rows = np.zeros(len(layers), dtype=dev.native.LAYER_DTYPE)
    keep = []
    for i, (img, pos, mode, pil, op) in enumerate(layers):
        t = dev.up(img)
        keep.append(t)
        p, w, h, s = dev.args(t)
        rows[i]["src"], rows[i]["src_w"], rows[i]["src_h"], rows[i]["src_stride"] = p, w, h, s
        rows[i]["M"] = [1, 0, 0, 0, 1, 0]
        rows[i]["bbox_x"], rows[i]["bbox_y"], rows[i]["bbox_w"], rows[i]["bbox_h"] = pos[0], pos[1], w, h
        rows[i]["blend_mode"], rows[i]["flags"], rows[i]["opacity"] = mode, 1 if pil else 0, op
    frame = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda") if initial is None else dev.up(initial)
    bg = np.asarray(bg_color, dtype=np.uint8)
    dev.native.check(dev.lib.mvb_composite_stack(frame.data_ptr(), W, H, frame.stride(0), bg.ctypes.data,
                                                 len(layers), rows.ctypes.data, flags, dev.stream()))
    return dev.down(frame)


def test_opaque_stack_every_mode_every_value_triple_vs_oracle(dev):
    """The fused kernel's opaque-frame instantiation (float arithmetic, stack_float.cuh), exhaustively:
    every (bg value, fg value, fg alpha) triple = 2^24 pixels per case, for the 18 numpy blend modes
    and PIL's over, at opacity 1 and at a fractional opacity; the layer overhangs the frame on every
    side so edge tiles are exercised too."""
    from oracle import oracle as O
    n = 4096
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    b = (yy & 255).astype(np.uint8)
    f = (xx & 255).astype(np.uint8)
    fa = ((xx >> 8) | ((yy >> 8) << 4)).astype(np.uint8)
    A = np.stack([b, b * 7 + 3, 255 - b, np.full_like(b, 255)], -1)
    B = np.stack([f, f * 5 + 1, f ^ 0x5A, fa], -1)
    pad = 5
    Bp = np.zeros((n + 2 * pad, n + 2 * pad, 4), np.uint8)
    Bp[pad:-pad, pad:-pad] = B
    Bp[:pad], Bp[-pad:], Bp[:, :pad], Bp[:, -pad:] = 77, 99, 11, 200      # overhang: never visible
    for mode in range(19):
        pil = mode == 18
        for op in (1.0, 0.7):
            got = _stack(dev, n, n, (9, 8, 7, 255), [(A, (0, 0), 0, True, 1.0), (Bp, (-pad, -pad), mode % 18, pil, op)])
            want = O.alpha_composite(A.copy(), B, (0, 0), op, mode % 18, 0, use_pil=pil)
            bad = np.nonzero((got != want).any(-1))
            assert len(bad[0]) == 0, (mode, op, len(bad[0]), got[bad][:3], want[bad][:3], A[bad][:3], B[bad][:3])


def test_translucent_stack_every_alpha_pair_every_mode_vs_oracle(dev):
    """The fused kernel's general instantiation (frame alpha carried as a fourth channel): every
    (frame alpha, layer alpha) pair with random colours, 18 numpy modes + PIL's over, opacity 1 and
    fractional, starting from given frame contents (MVB_STACK_KEEP_FRAME); the layer overhangs the
    frame so that pixels outside its bbox must stay untouched."""
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    n = 1024                                   # 4 x 4 repeats of the 256 x 256 alpha-pair grid
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    F = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    F[..., 3] = yy & 255
    F[:40, :40, :3] = [[0, 255, 128]]          # extremes
    B = rng.integers(0, 256, (n - 64, n - 32, 4), dtype=np.uint8)
    B[..., 3] = xx[: n - 64, : n - 32] & 255
    pos = (-9, 40)
    for mode in range(19):
        pil = mode == 18
        for op in (1.0, 0.37):
            got = _stack(dev, n, n, (0, 0, 0, 0), [(B, pos, mode % 18, pil, op)], flags=1, initial=F)
            want = O.alpha_composite(F.copy(), B, pos, op, mode % 18, 0, use_pil=pil)
            bad = np.nonzero((got != want).any(-1))
            assert len(bad[0]) == 0, (mode, op, len(bad[0]), got[bad][:3], want[bad][:3], F[bad][:3])


def test_alpha_composite_fuzz_vs_oracle(dev):
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    for i in range(120):
        bg = rng.integers(0, 256, tuple(rng.integers(1, 70, 2)) + (4,), dtype=np.uint8)
        fg = rng.integers(0, 256, tuple(rng.integers(1, 70, 2)) + (4,), dtype=np.uint8)
        if i % 2:
            bg[..., 3] = rng.choice([0, 1, 254, 255], size=bg.shape[:2])
        if i % 3 == 0:
            fg[..., 3] = rng.choice([0, 1, 254, 255], size=fg.shape[:2])
        pos = rng.integers(-60, 60, 2)
        op = [1.0, 0.0, float(rng.uniform())][i % 3]
        mode, matte = i % 18, (i // 18) % 3
        want = O.alpha_composite(bg, fg, pos, op, mode, matte)
        got = _composite(dev, bg, fg, pos, op, mode, matte)
        assert np.array_equal(got, want), (i, mode, matte)


def _warp(dev, src, M, dsize):
    import torch
    s = dev.up(src)
    d = torch.empty((dsize[1], dsize[0], 4), dtype=torch.uint8, device="cuda")
    sp, sw, sh, ss = dev.args(s)
    Mf = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(6))
    dev.native.check(dev.lib.mvb_warp_affine_rgba8(sp, sw, sh, ss, d.data_ptr(), dsize[0], dsize[1], d.stride(0),
                                                   Mf.ctypes.data, dev.stream()))
    return dev.down(d)


def test_warp_affine_golden(dev, golden):
    g = golden("warp_affine")
    for i, (M, dsize) in enumerate(zip(g["M"], g["dsize"])):
        out = _warp(dev, g[f"src{i}"], M, (int(dsize[0]), int(dsize[1])))
        assert np.array_equal(out, g[f"dst{i}"]), f"case {i}"


def test_warp_affine_large_vs_oracle(dev):
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (1080, 1920, 4), dtype=np.uint8)
    for ang, s, tx, ty in [(13.3, 1.17, 400.3, 100.7), (-170.0, 0.45, 1500.0, 900.25), (0.0, 1.0, 0.5, 0.5)]:
        a = np.deg2rad(ang)
        M = np.array([[s * np.cos(a), -s * np.sin(a), tx], [s * np.sin(a), s * np.cos(a), ty]])
        assert np.array_equal(_warp(dev, src, M, (2560, 1848)), O.warp_affine(src, M, (2560, 1848)))
    # a non-contiguous source view (ops.crop hands those out, ops.py:296): explicit row stride
    import torch
    big = dev.up(src)
    view = big[100:600, 200:900]
    d = torch.empty((300, 400, 4), dtype=torch.uint8, device="cuda")
    M = np.array([[0.7, 0.2, 10.0], [-0.2, 0.7, 50.0]])
    dev.native.check(dev.lib.mvb_warp_affine_rgba8(view.data_ptr(), 700, 500, view.stride(0), d.data_ptr(), 400, 300,
                                                   d.stride(0), np.ascontiguousarray(M.reshape(6)).ctypes.data,
                                                   dev.stream()))
    assert np.array_equal(dev.down(d), O.warp_affine(src[100:600, 200:900], M, (400, 300)))


def test_effects_golden(dev, golden):
    import movis_b200 as mv
    from movis_b200.contrib.segmentation import ChromaKey
    g = golden("effects")
    for i, p in enumerate(g["params"]):
        r, st, off, ang, op = (float(v) for v in p[:5])
        col = tuple(int(v) for v in p[5:8])
        hu, sat, lum = (float(v) for v in p[8:11])
        krange = tuple(float(v) for v in p[11:14])
        img = g[f"img{i}"]
        for name, eff, src in [
            ("blur", mv.effect.GaussianBlur(r), img), ("glow", mv.effect.Glow(r, st), img),
            ("shadow", mv.effect.DropShadow(r, off, ang, col, op), img), ("fill", mv.effect.FillColor(col), img),
            ("hsl", mv.effect.HSLShift(hu, sat, lum), img),
            ("chroma", ChromaKey(col, key_color_range=krange), g[f"keyed{i}"]),
        ]:
            got = eff(src, 0.0)
            assert isinstance(got, np.ndarray) and got.dtype == np.uint8
            assert got.shape == g[f"{name}{i}"].shape, (name, i)
            assert np.array_equal(got, g[f"{name}{i}"]), (name, i)
    blue = np.zeros((4, 4, 4), np.uint8)
    blue[..., 2] = 255
    blue[..., 3] = 255
    assert np.array_equal(mv.effect.HSLShift(hue=180.0)(blue, 0.0), g["hsl_blue_180"])


def test_blur_family_vs_oracle_all_radii(dev):
    import movis_b200 as mv
    from oracle import oracle as O
    rng = np.random.default_rng(9)
    for r in [0.3, 0.5, 1.0, 2.0, 3.3, 7.7, 10.0, 17.2, 25.0]:
        img = rng.integers(0, 256, tuple(rng.integers(8, 90, 2)) + (4,), dtype=np.uint8)
        assert np.array_equal(mv.effect.GaussianBlur(r)(img, 0.0), O.effect_gaussian_blur(img, r)), r
        assert np.array_equal(mv.effect.Glow(r, 1.3)(img, 0.0), O.effect_glow(img, r, 1.3)), r
        assert np.array_equal(mv.effect.DropShadow(r, 20.0, 45.0, (9, 8, 7), 0.5)(img, 0.0),
                              O.effect_drop_shadow(img, r, 20.0, 45.0, (9, 8, 7), 0.5)), r
    img = rng.integers(0, 256, (400, 400, 4), dtype=np.uint8)  # config-3 sized
    assert np.array_equal(mv.effect.GaussianBlur(10.0)(img, 0.0), O.effect_gaussian_blur(img, 10.0))


def test_pointwise_effects_vs_oracle_ragged_widths(dev):
    import movis_b200 as mv
    from movis_b200.contrib.segmentation import ChromaKey
    from oracle import oracle as O
    rng = np.random.default_rng(10)
    for w in (1, 31, 32, 33, 95, 100, 257):
        img = rng.integers(0, 256, (7, w, 4), dtype=np.uint8)
        img[:3, :, :3] = (0, 255, 0)
        assert np.array_equal(mv.effect.HSLShift(77.0, -0.2, 0.3)(img, 0.0), O.effect_hsl_shift(img, 77.0, -0.2, 0.3)), w
        assert np.array_equal(ChromaKey()(img, 0.0), O.effect_chroma_key(img)), w
    # strided inputs (crops of a wider device image: ops.crop hands those out), vector and scalar widths
    import torch
    from movis_b200.device import to_device
    wide = rng.integers(0, 256, (37, 300, 4), dtype=np.uint8)
    wide[5:20, 40:200, :3] = (0, 255, 0)
    dev = to_device(wide)
    for x0, ww in ((0, 300), (4, 256), (8, 100), (3, 97), (12, 4), (1, 1), (16, 960 // 4)):
        view = dev[2:35, x0:x0 + ww]
        host = np.ascontiguousarray(wide[2:35, x0:x0 + ww])
        got = ChromaKey().apply_device(view, 0.0).cpu().numpy()
        assert np.array_equal(got, O.effect_chroma_key(host)), (x0, ww)
        got = mv.effect.HSLShift(77.0, -0.2, 0.3).apply_device(view, 0.0).cpu().numpy()
        assert np.array_equal(got, O.effect_hsl_shift(host, 77.0, -0.2, 0.3)), (x0, ww)
        got = mv.effect.FillColor((9, 200, 30)).apply_device(view, 0.0).cpu().numpy()
        assert np.array_equal(got, O.effect_fill_color(host, (9, 200, 30))), (x0, ww)
    full = rng.integers(0, 256, (2160, 3840, 4), dtype=np.uint8)  # config-4 sized green screen
    full[:1080, :, :3] = (0, 255, 0)
    got = ChromaKey()(full, 0.0)
    assert np.array_equal(got, O.effect_chroma_key(full))
    assert (got[:1080, :, 3] == 0).all() and set(np.unique(got[..., 3])) <= {0, 255}


def test_chroma_key_every_rgb_colour_vs_oracle(dev):
    """The ChromaKey kernel tests integer bounds instead of converting to HSV (csrc/chroma_key.cuh): all
    2^24 colours for keys whose hue range is plain, wraps around red, is empty, is everything, and for
    grey / dark keys where cv2's division tables clamp."""
    from movis_b200.contrib.segmentation import ChromaKey
    from oracle import oracle as O
    v = np.arange(256, dtype=np.uint8)
    img = np.empty((4096, 4096, 4), dtype=np.uint8)
    img[..., 0] = np.tile(v, 16)[None, :]                       # r = x % 256
    img[..., 1] = np.repeat(v, 16)[:, None]                     # g = y // 16
    img[..., 2] = (np.arange(4096)[:, None] % 16) * 16 + np.arange(4096)[None, :] // 256   # b
    img[..., 3] = 77
    assert len(np.unique(img[..., :3].reshape(-1, 3).astype(np.uint32) @ np.array([1, 256, 65536], np.uint32))) == 1 << 24
    keys = [((0, 255, 0), (20.0, 0.3, 0.3)), ((255, 0, 0), (40.0, 0.5, 0.5)), ((250, 10, 30), (90.0, 0.2, 0.9)),
            ((0, 0, 255), (0.0, 0.0, 0.0)), ((10, 200, 220), (360.0, 1.0, 1.0)), ((128, 128, 128), (30.0, 0.1, 0.1)),
            ((3, 2, 1), (179.0, 0.6, 0.02)), ((255, 255, 0), (1.0, 0.01, 0.01))]
    for col, rng_ in keys:
        got = ChromaKey(col, key_color_range=rng_)(img, 0.0)
        want = O.effect_chroma_key(img, col, rng_)
        assert np.array_equal(got, want), (col, rng_, int((got[..., 3] != want[..., 3]).sum()))


def test_public_alpha_composite_semantics(dev):
    import torch
    import movis_b200 as mv
    from oracle import oracle as O
    rng = np.random.default_rng(2)
    bg = rng.integers(0, 256, (20, 30, 4), dtype=np.uint8)
    fg = rng.integers(0, 256, (10, 10, 4), dtype=np.uint8)
    keep = bg.copy()
    out = mv.alpha_composite(bg, fg, (5, 5), 0.5)                       # PIL path: new array
    assert np.array_equal(bg, keep) and out is not bg
    assert np.array_equal(out, O.alpha_composite(keep, fg, (5, 5), 0.5))
    out = mv.alpha_composite(bg, fg, (5, 5), 0.5, "screen")             # numpy path: in place
    assert out is bg and np.array_equal(bg, O.alpha_composite(keep, fg, (5, 5), 0.5, "screen"))
    ro = keep.copy()
    ro.flags.writeable = False
    out = mv.alpha_composite(ro, fg, (5, 5), 0.5, mv.BlendingMode.SCREEN)
    assert np.array_equal(ro, keep) and np.array_equal(out, bg)
    out = mv.alpha_composite(keep.copy(), fg, (5, 5), 0.5, "normal")     # str 'normal' -> numpy arithmetic
    assert np.array_equal(out, O.alpha_composite(keep, fg, (5, 5), 0.5, "normal", use_pil=False))
    with pytest.raises(ValueError):
        mv.alpha_composite(keep.copy(), fg, blending_mode="nope")
    with pytest.raises(AssertionError):
        mv.alpha_composite(keep[..., :3], fg)
    with pytest.raises(AssertionError):
        mv.alpha_composite(keep.copy(), fg, opacity=1.5)
    tb, tf = torch.from_numpy(keep).cuda(), torch.from_numpy(fg).cuda()
    out = mv.alpha_composite(tb, tf, (5, 5), 0.5, "multiply")
    assert out is tb and np.array_equal(out.cpu().numpy(), O.alpha_composite(keep, fg, (5, 5), 0.5, "multiply"))


def test_invalid_arguments_are_rejected(dev):
    import torch
    t = torch.zeros((8, 8, 4), dtype=torch.uint8, device="cuda")
    bg = np.zeros(4, np.uint8)
    lay = np.zeros(1, dtype=dev.native.LAYER_DTYPE)
    lay["bbox_w"], lay["bbox_h"], lay["src_w"], lay["src_h"], lay["src_stride"] = 4, 4, 4, 4, 16
    rc = dev.lib.mvb_composite_stack(t.data_ptr(), 8, 8, 32, bg.ctypes.data, 1, lay.ctypes.data, 0, dev.stream())
    assert rc == -1 and b"source" in dev.lib.mvb_last_error()           # null src
    lay["src"] = t.data_ptr()
    lay["blend_mode"] = 99
    assert dev.lib.mvb_composite_stack(t.data_ptr(), 8, 8, 32, bg.ctypes.data, 1, lay.ctypes.data, 0, dev.stream()) == -1
    assert dev.lib.mvb_composite_stack(t.data_ptr(), 8, 8, 30, bg.ctypes.data, 0, None, 0, dev.stream()) == -1
    assert dev.lib.mvb_alpha_composite(t.data_ptr(), 8, 8, 32, t.data_ptr(), 8, 8, 32, 0, 0, 2.0, 0, 0, 0, dev.stream()) == -1


def test_stripe_source_golden_and_vs_oracle(dev, golden):
    """mvb_source_stripe_rgba8 through movis_b200.layer.Stripe: the reference's images, then a seeded sweep
    (sizes up to 4K, every early-return branch) against the oracle."""
    import movis_b200 as mv
    from movis_b200 import scenes
    from oracle import oracle as O
    g = golden("stripe")
    for i, kw in enumerate(scenes.stripe_cases()):
        assert np.array_equal(mv.layer.Stripe(**kw)(0.0), g[f"img{i}"]), (i, kw)
    rng = np.random.default_rng(123)
    sizes = [(3840, 2160), (1921, 1079), (1, 517), (640, 1)] + [(int(rng.integers(1, 700)), int(rng.integers(1, 500))) for _ in range(20)]
    for size in sizes:
        kw = dict(angle=float(rng.uniform(-720, 720)), color1=tuple(int(v) for v in rng.integers(0, 256, 3)),
                  color2=tuple(int(v) for v in rng.integers(0, 256, 3)), opacity1=float(rng.uniform(0, 1)),
                  opacity2=float(rng.uniform(0, 1)), total_width=float(rng.uniform(0.3, 400)), phase=float(rng.uniform(-5, 5)),
                  ratio=float(rng.choice([0.0, 1.0, 0.004, 0.996, rng.uniform(0, 1), rng.uniform(0, 1), rng.uniform(0, 1)])))
        got = mv.layer.Stripe(size=size, **kw)(0.0)
        assert got.shape == (size[1], size[0], 4)
        assert np.array_equal(got, O.stripe(size, **kw)), (size, kw)
    assert mv.layer.Stripe(size=(8, 8), duration=1.0)(1.0) is None and mv.layer.Stripe(size=(8, 8))(-0.1) is None


# opacities whose float64 table trunc(a * opacity) is NOT floor(a * c) for any c (a few ulps below a
# fraction of small integers): k_stack_prep must detect them and fall back to the stored table
AWKWARD_OPACITIES = [0.19999999999999998, 0.39999999999999997, 0.7999999999999999, 0.8333333333333333,
                     0.09999999999999999, 0.988235294117647]


def test_stack_opacity_constant_and_table_fallback_vs_oracle(dev):
    """The fused kernel scales the sampled alpha by one fma per row pair (constant found per layer by
    k_stack_prep) instead of the reference's 256-entry table; every alpha value x many opacities, both
    arithmetic paths, opaque and translucent frames, incl. the opacities that need the table."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    n = 256
    fg = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    fg[..., 3] = np.arange(n, dtype=np.uint8)[None, :]
    base = rng.integers(0, 256, (n, n, 4), dtype=np.uint8)
    ops = AWKWARD_OPACITIES + [0.0, 1.0, 0.5, 1 / 255, 254 / 255, 0.2, 0.7, 1 / 3, 2 / 3, 0.999999, 1e-9]
    ops += [float(v) for v in rng.random(12)]
    for frame_alpha in (255, None):
        bgimg = base.copy()
        if frame_alpha is not None:
            bgimg[..., 3] = frame_alpha
        for mode, pil in ((0, True), (1, False), (11, False)):
            for op in ops:
                if frame_alpha is None:
                    got = _stack(dev, n, n, (0, 0, 0, 0), [(fg, (0, 0), mode, pil, op)], flags=1, initial=bgimg)
                else:
                    got = _stack(dev, n, n, (1, 2, 3, 255), [(bgimg, (0, 0), 0, True, 1.0), (fg, (0, 0), mode, pil, op)])
                want = O.alpha_composite(bgimg.copy(), fg, (0, 0), op, mode, 0, use_pil=pil)
                assert np.array_equal(got, want), (frame_alpha, mode, pil, op)


def _frames_call(dev, W, H, bg_color, per_frame_layers):
    """mvb_composite_frames over identity-warp layers: per frame a list of (tensor, (x, y), mode, pil, opacity)."""
    import torch
    n = len(per_frame_layers)
    offsets = np.zeros(n + 1, dtype=np.int32)
    offsets[1:] = np.cumsum([len(l) for l in per_frame_layers])
    rows = np.zeros(int(offsets[-1]), dtype=dev.native.LAYER_DTYPE)
    k = 0
    for layers in per_frame_layers:
        for (t, pos, mode, pil, op) in layers:
            p, w, h, s = dev.args(t)
            rows[k]["src"], rows[k]["src_w"], rows[k]["src_h"], rows[k]["src_stride"] = p, w, h, s
            rows[k]["M"] = [1, 0, 0, 0, 1, 0]
            rows[k]["bbox_x"], rows[k]["bbox_y"], rows[k]["bbox_w"], rows[k]["bbox_h"] = pos[0], pos[1], w, h
            rows[k]["blend_mode"], rows[k]["flags"], rows[k]["opacity"] = mode, 1 if pil else 0, op
            k += 1
    out = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    ptrs = (out.data_ptr() + out.stride(0) * np.arange(n, dtype=np.int64)).astype(np.uint64)
    bg = np.asarray(bg_color, dtype=np.uint8)
    dev.native.check(dev.lib.mvb_composite_frames(n, ptrs.ctypes.data, W, H, out.stride(1), bg.ctypes.data,
                                                  offsets.ctypes.data, rows.ctypes.data if len(rows) else None, 0, dev.stream()))
    return dev.down(out)


def test_frames_call_splits_runs_by_descriptors_and_long_stacks(dev):
    """One mvb_composite_frames call whose frames need more distinct TMA descriptors than a launch
    carries (every layer has its own source image), with empty frames and a 40-layer frame in the
    middle: every frame must equal the same stack rendered alone through mvb_composite_stack."""
    import torch
    rng = np.random.default_rng(9)
    W, H = 96, 80
    pool = torch.from_numpy(rng.integers(0, 256, (260, 48, 64, 4), dtype=np.uint8)).cuda()
    per_frame, k = [], 0
    for f in range(44):
        n_layers = 40 if f == 20 else (0 if f in (3, 43) else 5)
        layers = []
        for i in range(n_layers):
            pos = (int(rng.integers(-20, W - 10)), int(rng.integers(-20, H - 10)))
            mode = int(rng.integers(0, 18))
            layers.append((pool[k % 260], pos, mode, mode == 0, float(rng.choice([1.0, 0.6, 0.25]))))
            k += 1
        per_frame.append(layers)
    for bg in ((5, 6, 7, 255), (0, 0, 0, 0)):
        got = _frames_call(dev, W, H, bg, per_frame)
        for f, layers in enumerate(per_frame):
            want = _frames_call(dev, W, H, bg, [layers])[0]
            assert np.array_equal(got[f], want), (bg, f)
        # and the single-frame path against the oracle for a few of them
        from oracle import oracle as O
        for f in (0, 20, 43):
            ref = np.empty((H, W, 4), np.uint8)
            ref[...] = np.asarray(bg, np.uint8)
            for (t, pos, mode, pil, op) in per_frame[f]:
                ref = O.alpha_composite(ref, t.cpu().numpy(), pos, op, mode, 0, use_pil=pil)
            assert np.array_equal(got[f], ref), (bg, f)

"""Generate tests/golden/*.npz from the UNMODIFIED reference (rezoo/movis at /root/reference).

Run in the build container (the reference does not travel to the GPU box):

    python -m oracle.gen_golden

Everything stored here is an output of the reference's own code (numpy / cv2 / PIL underneath) on
seeded inputs; the inputs are either stored alongside or regenerated from the seed by the tests
through ``movis_b200.scenes``.  The oracle (oracle/mvo.c) and the CUDA path are both tested
against these files.  Versions used for the committed vectors are recorded in manifest.json.
"""
from __future__ import annotations

import hashlib
import json
import os
import warnings

import numpy as np

from movis_b200 import scenes
from oracle.ref_harness import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def save(name: str, **arrays) -> None:
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(f"  {name}.npz  {os.path.getsize(os.path.join(OUT, name + '.npz')) / 1e3:.0f} kB")


def gen_stripe(mv) -> None:
    """Stripe layers (layer/texture.py:88-191): stand-alone images for scenes.stripe_cases() and frames
    of scenes.build_stripes (Stripe sources under animated transforms / blend modes)."""
    out = {}
    for i, kw in enumerate(scenes.stripe_cases()):
        out[f"img{i}"] = mv.layer.Stripe(**kw)(0.0)
    comp = scenes.build_stripes(mv)
    times = [0.0, 1.37, 3.9]
    out["times"] = np.asarray(times)
    for t in times:
        out[f"frame_{t:.3f}"] = comp(t)
    comp2 = scenes.build_stripes(mv, div=2)
    comp2.preview_level = 2 if hasattr(comp2, "preview_level") else 1
    out["frame_div2_L2"] = comp2(2.2)
    save("stripe", **out)


def gen_ops(mv) -> None:
    """ops.crop + AlphaMatte / LuminanceMatte under transforms (scenes.build_ops), frames of the reference."""
    out = {}
    comp = scenes.build_ops(mv)
    times = [0.0, 0.9, 1.7]
    out["times"] = np.asarray(times)
    for t in times:
        out[f"frame_{t:.3f}"] = comp(t)
    out["frame_opaque_0.900"] = comp(0.9, bg_color=(5, 6, 7, 255))
    comp2 = scenes.build_ops(mv, div=2)
    with comp2.preview(level=2):
        out["frame_div2_L2"] = comp2(1.3)
    # the layers on their own (strided views / matte images as the reference hands them to the stack)
    for name in ("crop", "alpha_matte", "lum_matte", "crop_comp"):
        out[f"layer_{name}"] = np.ascontiguousarray(comp[name].layer(0.9))
    save("ops", **out)


def gen_full_size(mv) -> None:
    """sha256 of one FULL-SIZE frame of every BASELINE.json config, rendered by the reference itself
    (scenes.full_size_shots; 1 - 20 s per frame).  Only the digests are stored: the tests render the
    same frame on the GPU and compare hashes, so the frames need not be committed."""
    import time
    from movis.contrib.segmentation import ChromaKey
    path = os.path.join(OUT, "manifest.json")
    with open(path) as fh:
        manifest = json.load(fh)
    out = manifest.setdefault("full_size_digests", {})
    for tag, build, t in scenes.full_size_shots(mv, ChromaKey):
        t0 = time.perf_counter()
        frame = np.asarray(build()(t, bg_color=(0, 0, 0, 255)))
        out[tag] = {"time": t, "shape": list(frame.shape), "sha256": sha(frame)}
        print(f"  {tag}: {frame.shape} {time.perf_counter() - t0:.1f} s  {out[tag]['sha256'][:16]}")
    with open(path, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


def main() -> None:
    warnings.filterwarnings("ignore")
    os.makedirs(OUT, exist_ok=True)
    mv = import_reference()
    import sys
    if "fullsize" in sys.argv[1:]:   # only the full-size digests (merged into manifest.json)
        gen_full_size(mv)
        return
    if "stripe" in sys.argv[1:]:   # only this file (the others are unchanged)
        gen_stripe(mv)
        return
    if "ops" in sys.argv[1:]:
        gen_ops(mv)
        return
    import cv2
    import PIL
    from movis.contrib.segmentation import ChromaKey
    from movis.enum import BlendingMode, MatteMode
    from movis.imgproc import BLENDING_MODE_TO_FUNC
    from movis.layer.composition import _get_fixed_affine_matrix

    # --- 1. blend-mode tables T[b, f] as the reference functions produce them inside _overlay ---
    b = np.repeat(np.arange(256, dtype=np.uint32)[:, None], 256, 1)[..., None]
    f = np.repeat(np.arange(256, dtype=np.uint32)[None, :], 256, 0)[..., None]
    luts = []
    for m in BlendingMode:
        with np.errstate(all="ignore"):
            T = BLENDING_MODE_TO_FUNC[m](b, f)
        c1 = np.full_like(b, 65025)
        luts.append(((c1 * T) // c1).astype(np.uint8)[..., 0])
    save("blend_luts", luts=np.stack(luts))

    # --- 2. alpha_composite: every mode x matte on edge-case alphas, positions in/out of frame ---
    rng = np.random.default_rng(11)
    cases = {}
    meta = []
    for i in range(72):
        bh, bw = (int(v) for v in rng.integers(3, 40, 2))
        fh, fw = (int(v) for v in rng.integers(3, 40, 2))
        bg = rng.integers(0, 256, (bh, bw, 4), dtype=np.uint8)
        fg = rng.integers(0, 256, (fh, fw, 4), dtype=np.uint8)
        if i % 2:
            bg[..., 3] = rng.choice([0, 1, 127, 128, 254, 255], size=(bh, bw))
        if i % 3:
            fg[..., 3] = rng.choice([0, 1, 127, 128, 254, 255], size=(fh, fw))
        if i % 7 == 0:
            bg[..., 3] = 255
        pos = (int(rng.integers(-20, 30)), int(rng.integers(-20, 30)))
        if i == 5:
            pos = (1000, 1000)  # entirely outside (tests/test_imgproc.py:41-50)
        op = [1.0, 0.0, 0.5, float(rng.uniform()), float(rng.uniform())][i % 5]
        mode = list(BlendingMode)[i % 18]
        matte = list(MatteMode)[(i // 18) % 3]
        out = mv.alpha_composite(bg.copy(), fg, pos, op, mode, matte)
        cases[f"bg{i}"], cases[f"fg{i}"], cases[f"out{i}"] = bg, fg, np.asarray(out)
        meta.append([pos[0], pos[1], mode.value, matte.value])
        cases[f"op{i}"] = np.float64(op)
    cases["meta"] = np.array(meta, dtype=np.int32)
    # string 'normal' goes down the numpy path (imgproc.py:270-273)
    bg = rng.integers(0, 256, (24, 31, 4), dtype=np.uint8)
    fg = rng.integers(0, 256, (17, 20, 4), dtype=np.uint8)
    cases["str_bg"], cases["str_fg"] = bg, fg
    cases["str_out"] = mv.alpha_composite(bg.copy(), fg, (3, 4), 0.6, "normal")
    save("alpha_composite", **cases)

    # exhaustive (fg alpha, bg alpha) pairs for both over operators
    sa, da = np.meshgrid(np.arange(256), np.arange(256))
    bg = rng.integers(0, 256, (256, 256, 4), dtype=np.uint8)
    fg = rng.integers(0, 256, (256, 256, 4), dtype=np.uint8)
    bg[..., 3], fg[..., 3] = da, sa
    save("over_alpha_pairs", bg=bg, fg=fg,
         pil_1=np.asarray(mv.alpha_composite(bg.copy(), fg, (0, 0), 1.0)),
         pil_037=np.asarray(mv.alpha_composite(bg.copy(), fg, (0, 0), 0.37)),
         soft_1=mv.alpha_composite(bg.copy(), fg, (0, 0), 1.0, BlendingMode.SOFT_LIGHT),
         mult_037=mv.alpha_composite(bg.copy(), fg, (0, 0), 0.37, BlendingMode.MULTIPLY))

    # --- 3. cv2.warpAffine on small sources with assorted matrices ------------------------------
    rng = np.random.default_rng(12)
    cases = {}
    mats, sizes = [], []
    for i in range(40):
        sh, sw = (int(v) for v in rng.integers(1, 48, 2))
        src = rng.integers(0, 256, (sh, sw, 4), dtype=np.uint8)
        ang, s = rng.uniform(-np.pi, np.pi), rng.uniform(0.2, 3, 2)
        M = np.array([[s[0] * np.cos(ang), -s[0] * np.sin(ang), rng.uniform(-10, 40)],
                      [s[1] * np.sin(ang), s[1] * np.cos(ang), rng.uniform(-10, 40)]])
        if i % 7 == 0:
            M = np.array([[1, 0, float(rng.integers(-5, 5))], [0, 1, float(rng.integers(-5, 5))]])
        if i % 11 == 0:
            M = np.array([[1, 0, rng.uniform(-5, 5)], [0, 1, rng.uniform(-5, 5)]])
        if i == 39:
            M = np.zeros((2, 3))  # singular: every pixel samples (0, 0)
        dw, dh = (int(v) for v in rng.integers(1, 64, 2))
        cases[f"src{i}"] = src
        cases[f"dst{i}"] = cv2.warpAffine(src, M, dsize=(dw, dh), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        mats.append(M.reshape(6))
        sizes.append([dw, dh])
    cases["M"] = np.array(mats)
    cases["dsize"] = np.array(sizes, dtype=np.int32)
    save("warp_affine", **cases)

    # --- 4. Gaussian: Q0.8 weights recovered through cv2 + blurred images -----------------------
    rng = np.random.default_rng(13)
    cases = {}
    radii = [0.3, 0.5, 0.77, 1.0, 1.5, 2.0, 3.0, 3.3, 5.0, 7.7, 10.0, 13.0, 17.2, 25.0]
    cases["radii"] = np.array(radii)
    for i, r in enumerate(radii):
        k = 2 * int(np.ceil(max(1, (4 * r - 1) / 2))) + 1
        src4 = rng.integers(0, 256, (23, 31, 4), dtype=np.uint8)
        src1 = rng.integers(0, 256, (19, 45), dtype=np.uint8)
        cases[f"src4_{i}"], cases[f"src1_{i}"] = src4, src1
        cases[f"dst4_{i}"] = cv2.GaussianBlur(src4, (k, k), sigmaX=r, sigmaY=r)
        cases[f"dst1_{i}"] = cv2.GaussianBlur(src1, (k, k), sigmaX=r, sigmaY=r)
    save("gaussian_blur", **cases)

    # --- 5. effects ---------------------------------------------------------------------------
    rng = np.random.default_rng(14)
    cases = {}
    params = []
    for i in range(12):
        h, w = (int(v) for v in rng.integers(5, 50, 2))
        img = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        if i % 2:
            img[..., 3] = rng.choice([0, 1, 128, 255], size=(h, w))
        r = [0.0, 0.4, 1.0, 2.5, 6.0, 10.0][i % 6]
        st = float(rng.uniform(0, 3))
        off, ang, op = float(rng.uniform(0, 20)), float(rng.uniform(-180, 360)), float(rng.uniform())
        col = [int(v) for v in rng.integers(0, 256, 3)]
        hu, sat, lum = float(rng.uniform(-400, 400)), float(rng.uniform(-1, 1)), float(rng.uniform(-1, 1))
        cases[f"img{i}"] = img
        cases[f"blur{i}"] = mv.effect.GaussianBlur(r)(img, 0.0)
        cases[f"glow{i}"] = mv.effect.Glow(r, st)(img, 0.0)
        cases[f"shadow{i}"] = mv.effect.DropShadow(r, off, ang, tuple(col), op)(img, 0.0)
        cases[f"fill{i}"] = mv.effect.FillColor(tuple(col))(img, 0.0)
        cases[f"hsl{i}"] = mv.effect.HSLShift(hu, sat, lum)(img, 0.0)
        keyed = img.copy()
        keyed[: h // 2, :, :3] = col
        krange = (float(rng.uniform(0, 90)), float(rng.uniform()), float(rng.uniform()))
        cases[f"keyed{i}"] = keyed
        cases[f"chroma{i}"] = ChromaKey(tuple(col), key_color_range=krange)(keyed, 0.0)
        params.append([r, st, off, ang, op, *col, hu, sat, lum, *krange])
    cases["params"] = np.array(params)
    e = ChromaKey()
    cases["chroma_default_bounds"] = np.stack([e._lower_color, e._upper_color])
    # pure-colour HSLShift checks of tests/effect/test_color.py:27-33
    blue = np.zeros((4, 4, 4), np.uint8)
    blue[..., 2] = 255
    blue[..., 3] = 255
    cases["hsl_blue_180"] = mv.effect.HSLShift(hue=180.0)(blue, 0.0)
    save("effects", **cases)

    # --- 6. colour conversions on a stratified sample ------------------------------------------
    rng = np.random.default_rng(15)
    rgb = rng.integers(0, 256, (64, 1000, 3), dtype=np.uint8)
    rgb[0, :256] = np.stack([np.arange(256)] * 3, -1)  # greys
    hsv_in = rng.integers(0, 256, (64, 1000, 3), dtype=np.uint8)
    hsv_in[..., 0] = rng.integers(0, 181, (64, 1000))
    save("color_convert", rgb=rgb, hsv=cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV), hsv_in=hsv_in,
         rgb_out=cv2.cvtColor(hsv_in, cv2.COLOR_HSV2RGB))

    # --- 7. host engine: attribute values and affine plans ---------------------------------------
    rng = np.random.default_rng(16)
    easings = [e.name.lower() for e in mv.Easing]
    cases = {"easings": np.array(easings)}
    ts = np.concatenate([np.linspace(-0.5, 5.5, 61), np.arange(0, 5, 1 / 30)[:30]])
    cases["times"] = ts
    vals = []
    for e in easings:
        a = mv.Attribute(0.0, mv.AttributeType.SCALAR)
        a.enable_motion().extend([0.0, 2.0, 5.0], [1.0, -3.5, 7.25], [e, e])
        vals.append(np.array([a(t)[0] for t in ts]))
    cases["easing_values"] = np.stack(vals)
    tv, plans = [], []
    dirs = list(mv.Direction)
    for i in range(400):
        w, h = (int(v) for v in rng.integers(1, 2000, 2))
        generic = i % 5 != 0
        rot = float(rng.uniform(-720, 720)) if generic else float(rng.choice([0, 90, 180, 270, 45, -90]))
        pos = (float(rng.uniform(-500, 4000)), float(rng.uniform(-500, 2500))) if generic else \
            (float(rng.integers(0, 4000)), float(rng.integers(0, 2000)))
        sc = (float(rng.uniform(0.01, 4)), float(rng.uniform(0.01, 4))) if generic else (1.0, 1.0)
        anc = (float(rng.uniform(-9, 9)), float(rng.uniform(-9, 9))) if generic else (0.0, 0.0)
        d = dirs[int(rng.integers(0, 9))]
        L = int(rng.choice([1, 2, 3, 4]))
        p = mv.transform.TransformValue(anc, pos, sc, rot, 1.0, d, mv.BlendingMode.NORMAL)
        r = _get_fixed_affine_matrix(np.empty((h, w, 4), np.uint8), p, L)
        tv.append([w, h, *anc, *pos, *sc, rot, d.value, L])
        if r is None:
            plans.append([0] * 6 + [0, 0, 0, 0, 0])
        else:
            A, (W, H), (ox, oy) = r
            plans.append([*A.reshape(6), ox, oy, int(W), int(H), 1])
    cases["transform_inputs"] = np.array(tv)
    cases["affine_plans"] = np.array(plans)
    save("host_engine", **cases)

    # --- 8. whole-scene frames -----------------------------------------------------------------
    frames = {}
    digests = {}

    def shoot(tag, comp, times, bg, levels=(1,), store=True):
        for L in levels:
            for t in times:
                with comp.preview(level=L):
                    fr = np.asarray(comp(t, bg_color=bg))
                key = f"{tag}_L{L}_t{t:.3f}"
                digests[key] = sha(fr)
                if store:
                    frames[key] = fr

    for seed in range(4):
        shoot(f"fuzz{seed}", scenes.build_fuzz(mv, seed), [0.0, 0.7, 1.9],
              (0, 0, 0, 0) if seed % 2 else (9, 8, 7, 255), levels=(1, 2, 4))
    shoot("s1", scenes.build_s1(mv, div=12), [0.0, 1.3, 4.9], (0, 0, 0, 255))
    shoot("s2v0", scenes.build_s2(mv, div=12, variant=0), [0.0, 2.1, 4.9], (0, 0, 0, 255), levels=(1, 2))
    shoot("s2v2", scenes.build_s2(mv, div=12, variant=2), [0.5, 3.3], (0, 0, 0, 255))
    shoot("s2v0t", scenes.build_s2(mv, div=12, variant=0), [1.0], (0, 0, 0, 0))  # transparent frame
    shoot("s3", scenes.build_s3(mv, div=6), [0.0, 2.5], (0, 0, 0, 255))
    shoot("s4", scenes.build_s4(mv, ChromaKey, div=12), [0.0, 1.01], (0, 0, 0, 255), levels=(1, 2, 4))
    shoot("s5", scenes.build_s5(mv, ChromaKey, div=12, duration=6.0), [0.0, 3.1, 5.9], (0, 0, 0, 255), levels=(1, 2))
    # digests only (frames regenerated by the tests): denser time sampling of the headline scene
    shoot("s2v0d", scenes.build_s2(mv, div=12, variant=0), [round(0.1 + 0.4 * i, 3) for i in range(12)],
          (0, 0, 0, 255), store=False)
    save("scene_frames", **frames)
    gen_stripe(mv)
    gen_ops(mv)
    try:
        with open(os.path.join(OUT, "manifest.json")) as fh:
            previous = json.load(fh)
    except Exception:
        previous = {}
    manifest = {
        "full_size_digests": previous.get("full_size_digests", {}),
        "reference": "rezoo/movis 0.7.1 (/root/reference, unmodified; Qt/imageio/audio modules stubbed)",
        "numpy": np.__version__, "opencv": cv2.__version__, "pillow": PIL.__version__,
        "digests": digests,
    }
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("done")


if __name__ == "__main__":
    main()

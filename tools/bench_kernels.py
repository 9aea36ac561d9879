"""Kernel-level timings of the stand-alone operators at 4K through the C ABI with PREALLOCATED buffers
(no allocator / Python object work between launches), CUDA events around `reps` back-to-back calls on
two alternating input images (66 MB apart, so no call finds its input in L2 from the call before... the
126 MB L2 still holds part of it: the GB/s are upper bounds of the HBM rate, the times are what a frame pays).

    python tools/bench_kernels.py [reps]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movis_b200 import _native  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
L = _native.lib()
H, W = 2160, 3840
g = torch.Generator(device="cuda").manual_seed(0)
src = [torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
for s in src:
    s[: H // 2, :, 0] = 0
    s[: H // 2, :, 1] = 255
    s[: H // 2, :, 2] = 0
dst = torch.empty((H + 200, W + 200, 4), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def timed(name, fn, nbytes):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name:34s} {us:8.1f} us   {nbytes / us / 1e3:8.1f} GB/s")


lo = np.array([50, 178, 178], np.uint8)
hi = np.array([70, 255, 255], np.uint8)
timed("chroma_key", lambda i: _native.check(L.mvb_effect_chroma_key_rgba8(
    src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), lo.ctypes.data, hi.ctypes.data, stream)), 2 * H * W * 4)
col = np.array([255, 0, 0], np.uint8)
timed("fill_color", lambda i: _native.check(L.mvb_effect_fill_color_rgba8(
    src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), col.ctypes.data, stream)), 2 * H * W * 4)
x = np.arange(256, dtype=np.uint8)
luts = [np.ascontiguousarray((x.astype(np.int32) * 3 // 4).astype(np.uint8)) for _ in range(3)]
timed("hsl_shift", lambda i: _native.check(L.mvb_effect_hsl_shift_rgba8(
    src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), luts[0].ctypes.data, luts[1].ctypes.data,
    luts[2].ctypes.data, stream)), 2 * H * W * 4)
for radius in (3.0, 10.0):
    k = _native.blur_ksize(radius)
    wts = _native.gaussian_weights_q8(radius, k)
    scratch = torch.empty(L.mvb_effect_blur_scratch_bytes(W, H, k, 4), dtype=torch.uint8, device="cuda")
    glow = np.clip(1.3 * x.astype(np.float32), 0, 255).astype(np.uint8)
    out_bytes = (H + 2 * k) * (W + 2 * k) * 4
    timed(f"gaussian_blur r={radius:g} (k={k})", lambda i: _native.check(L.mvb_effect_blur_rgba8(
        src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), k, wts.ctypes.data, _native.MVB_BLUR_PLAIN, None,
        scratch.data_ptr(), stream)), H * W * 4 + out_bytes)
    timed(f"glow r={radius:g}", lambda i: _native.check(L.mvb_effect_blur_rgba8(
        src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), k, wts.ctypes.data, _native.MVB_BLUR_GLOW,
        glow.ctypes.data, scratch.data_ptr(), stream)), H * W * 4 + out_bytes)
bg = torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g)
for mode, flags, name in ((0, _native.MVB_LAYER_PIL_OVER, "alpha_composite normal (PIL)"), (3, 0, "alpha_composite overlay")):
    timed(name, lambda i: _native.check(L.mvb_alpha_composite(
        bg.data_ptr(), W, H, W * 4, src[i & 3].data_ptr(), W, H, W * 4, 5, 7, 0.8, mode, 0, flags, stream)), 3 * H * W * 4)

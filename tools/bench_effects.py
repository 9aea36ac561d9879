"""Device timings of the stand-alone operators at 4K (3840x2160 RGBA), CUDA events, 20 repetitions:

    python tools/bench_effects.py

Prints microseconds per call and algorithmic GB/s (input image + output image, each counted once)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402
from movis_b200.device import to_device  # noqa: E402
from movis_b200 import imgproc  # noqa: E402

rng = np.random.default_rng(0)
H, W = 2160, 3840
img = to_device(rng.integers(0, 256, (H, W, 4), dtype=np.uint8))
bg = to_device(rng.integers(0, 256, (H, W, 4), dtype=np.uint8))


def timed(fn, reps=20):
    out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, out


cases = [
    ("GaussianBlur(10)", lambda: mv.effect.GaussianBlur(10.0).apply_device(img, 0.0)),
    ("GaussianBlur(3)", lambda: mv.effect.GaussianBlur(3.0).apply_device(img, 0.0)),
    ("Glow(10, 1.3)", lambda: mv.effect.Glow(10.0, 1.3).apply_device(img, 0.0)),
    ("DropShadow(10, 20, 45, 0.5)", lambda: mv.effect.DropShadow(radius=10.0, offset=20.0, angle=45.0, opacity=0.5).apply_device(img, 0.0)),
    ("ChromaKey()", lambda: ChromaKey().apply_device(img, 0.0)),
    ("FillColor", lambda: mv.effect.FillColor(color=(255, 0, 0)).apply_device(img, 0.0)),
    ("HSLShift(hue=40)", lambda: mv.effect.HSLShift(hue=40.0).apply_device(img, 0.0)),
    ("alpha_composite normal (PIL)", lambda: imgproc.alpha_composite(bg, img, (5, 7), 0.8, "normal")),
    ("alpha_composite overlay", lambda: imgproc.alpha_composite(bg, img, (5, 7), 0.8, "overlay")),
]
x = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
for _ in range(50):          # bring the clocks up before the first measurement
    x.add_(1)
torch.cuda.synchronize()
for name, fn in cases + cases[:1]:
    us, out = timed(fn)
    nbytes = img.numel() + out.numel()
    print(f"{name:32s} {us:9.1f} us   {nbytes / us / 1e3:8.1f} GB/s   out {tuple(out.shape)}")

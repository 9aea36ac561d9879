"""Device-side frame rates of the other BASELINE.json configs (parity-test scenes, not bench lines):

    python tools/bench_scenes.py [scene ...]      (s1 s2 s2d s3 s4 s4L2 s4L4; default: all)

Each scene's 5 s @ 30 fps timeline is rendered through the public batched path
(Composition.render_frames); the second pass is timed with CUDA events, so per-frame host work
(effect kernels for animated effects, planning) is inside the number, source uploads are not.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402

BUILD = {
    "s1": (lambda: scenes.build_s1(mv), 1),
    "s2": (lambda: scenes.build_s2(mv), 1),
    "s2d": (lambda: scenes.build_s2d(mv), 1),
    "s3": (lambda: scenes.build_s3(mv), 1),
    "s4": (lambda: scenes.build_s4(mv, ChromaKey), 1),
    "s4L2": (lambda: scenes.build_s4(mv, ChromaKey), 2),
    "s4L4": (lambda: scenes.build_s4(mv, ChromaKey), 4),
}

for name in (sys.argv[1:] or list(BUILD)):
    build, level = BUILD[name]
    comp = build()
    times = np.arange(0, 5, 1 / 30.0)
    if name == "s2d":
        times = times[:60]          # 17 full-frame 4K layers: keep the batch at 2 GB
    with comp.preview(level=level):
        out = comp.render_frames(times)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        out = comp.render_frames(times, out=out)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"{name:5s} {out.shape[2]}x{out.shape[1]}  {len(times)} frames  {ms / len(times) * 1e3:8.1f} us/frame  {len(times) / ms * 1e3:8.0f} frames/s")
    del out, comp
    torch.cuda.empty_cache()

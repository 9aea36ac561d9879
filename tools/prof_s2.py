"""Short S2 run for ncu: renders a few full-size frames through the public batched path.

    python tools/prof_s2.py [n_frames] [scene]     (scene: s2 | s2d | s1 | s3 | s4 | s5)
"""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
scene = sys.argv[2] if len(sys.argv) > 2 else "s2"
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402
builders = {"s2": scenes.build_s2, "s2d": scenes.build_s2d, "s1": scenes.build_s1, "s3": scenes.build_s3,
            "s4": lambda m, div: scenes.build_s4(m, ChromaKey, div=div),
            "s5": lambda m, div: scenes.build_s5(m, ChromaKey, div=div)}
comp = builders[scene](mv, div=1)
times = (np.arange(n) * 7 % 150) / 30.0
out = comp.render_frames(times)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
out = comp.render_frames(times, out=out)
ev1.record()
torch.cuda.synchronize()
print(f"{scene}: {n} frames, {ev0.elapsed_time(ev1) / n * 1e3:.1f} us/frame, checksum {int(out.sum())}")

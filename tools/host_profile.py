"""cProfile of the host side of Composition.render_frames (plan_range + launch_plan) per scene."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "s1"
comp = {"s1": lambda: scenes.build_s1(mv), "s2": lambda: scenes.build_s2(mv), "s3": lambda: scenes.build_s3(mv),
        "s4": lambda: scenes.build_s4(mv, ChromaKey), "s5": lambda: scenes.build_s5(mv, ChromaKey)}[which]()
n = 150
times = np.arange(n) / 30.0
out = comp.render_frames(times)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = comp.render_frames(times, out=out)
host = time.perf_counter() - t0
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(f"{which}: host {host / n * 1e6:.1f} us/frame, host+device {total / n * 1e6:.1f} us/frame")
pr = cProfile.Profile()
pr.enable()
out = comp.render_frames(times, out=out)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats(sys.argv[2] if len(sys.argv) > 2 else "cumulative").print_stats(28)

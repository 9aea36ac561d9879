"""A/B timing of the fused kernel on full-size S2 (or s1/s2d/s4/s5): plan once, launch the whole timeline
`reps` times, CUDA events on the launch stream.  Kernel variants are selected by environment variables
read by the library at first use (MVB_ROWS, MVB_RUN_FRAMES, MVB_STAGES_LOG2), so run one process per variant:
    MVB_ROWS=8 python tools/ab_s2.py s2 3
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402
from movis_b200.render import launch_plan, plan_range  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "s2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 150
if which == "s2":
    comp, nbytes = scenes.build_s2(mv), scenes.s2_bytes_per_frame()
elif which == "s2d":
    comp, nbytes = scenes.build_s2d(mv), scenes.s2d_bytes_per_frame()
    n = 30
elif which == "s1":
    comp, nbytes = scenes.build_s1(mv), scenes.s1_bytes_per_frame()
elif which == "s4":
    comp, nbytes = scenes.build_s4(mv, ChromaKey), 4 * 3840 * 2160 + 2 * 33177600
    n = 48
elif which == "s3":
    comp, nbytes = scenes.build_s3(mv), 0
    n = 60
elif which == "s5":
    comp, nbytes = scenes.build_s5(mv, ChromaKey), 276614192
    n = 60
else:
    raise SystemExit("unknown scene")
times = np.arange(n) / 30.0
if which == "s5":
    times = times * 30.0   # spread over the 60 s timeline
H, W = comp.frame_shape()
frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
bg = (0, 0, 0, 255)
t0 = time.perf_counter()
plans = [plan_range(comp, times[i:i + 30]) for i in range(0, n, 30)]
plan_s = time.perf_counter() - t0


def go():
    for k, p in enumerate(plans):
        launch_plan(p, frames[30 * k:30 * k + 30], bg)


go()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    go()
e1.record()
host_s = time.perf_counter() - t0
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * n)
env = {k: v for k, v in os.environ.items() if k.startswith("MVB_")}
print(f"{which} {env}: {us:.1f} us/frame device, {nbytes / us / 1e3:.0f} GB/s algorithmic, "
      f"host enqueue {host_s / (reps * n) * 1e6:.1f} us/frame, plan {plan_s / n * 1e6:.1f} us/frame, "
      f"checksum {int(frames[::37].sum().item())}")

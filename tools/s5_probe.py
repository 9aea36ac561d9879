"""Per-call device and host times of the S5 timeline (1800 frames in 240-frame render_frames calls)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv
from movis_b200 import scenes
from movis_b200.contrib.segmentation import ChromaKey
from movis_b200.render import render_frames
comp = scenes.build_s5(mv, ChromaKey)
times = np.arange(1800) / 30.0
CALL = 240
out = torch.empty((CALL,) + comp.frame_shape() + (4,), dtype=torch.uint8, device="cuda")
for _ in range(2):
    render_frames(comp, times[:CALL], bg_color=(0, 0, 0, 255), out=out)
torch.cuda.synchronize()
for rep in range(2):
    t_all = time.perf_counter()
    for i in range(0, 1800, CALL):
        chunk = times[i:i + CALL]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        render_frames(comp, chunk, bg_color=(0, 0, 0, 255), out=out[:len(chunk)])
        e1.record(); host = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"rep {rep} frames {i:4d}..: host {host / len(chunk) * 1e6:6.1f} us/frame, device {e0.elapsed_time(e1) / len(chunk) * 1e3:6.1f} us/frame")
    print(f"rep {rep}: {1800 / (time.perf_counter() - t_all):.0f} frames/s (synchronised per call)")

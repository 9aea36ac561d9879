"""Where do the ~40 us per frame between ncu's kernel duration and the back-to-back rate go?"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv
from movis_b200 import scenes, _native
from movis_b200.render import plan_range, launch_plan

comp = scenes.build_s2(mv, div=1)
times = (np.arange(150) % 150) / 30.0
H, W = comp.frame_shape()
frames = torch.empty((150, H, W, 4), dtype=torch.uint8, device="cuda")
plan = plan_range(comp, times)
for _ in range(2):
    launch_plan(plan, frames, (0, 0, 0, 255))
torch.cuda.synchronize()

def timed(fn, n=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3 / 150, t_host / n * 1e6 / 150

print("150 frames in one C call: %.1f us/frame device, %.1f us/frame host enqueue" % timed(lambda: launch_plan(plan, frames, (0, 0, 0, 255))))
# per-frame events
lib = _native.lib()
bg = np.array([0, 0, 0, 255], np.uint8)
evs = [torch.cuda.Event(enable_timing=True) for _ in range(151)]
evs[0].record()
for f in range(150):
    d = plan.descriptors[plan.offsets[f]:plan.offsets[f + 1]]
    lib.mvb_composite_stack(frames[f].data_ptr(), W, H, frames.stride(1), bg.ctypes.data, len(d), d.ctypes.data, 0, torch.cuda.current_stream().cuda_stream)
    evs[f + 1].record()
torch.cuda.synchronize()
per = np.array([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(150)])
print("per-frame kernel times (event to event): mean %.1f  min %.1f  max %.1f us; total %.1f us/frame" % (per.mean(), per.min(), per.max(), evs[0].elapsed_time(evs[150]) * 1e3 / 150))
print("by frame index:", np.round(per[::15], 1))
# same frame repeated 150 times (no variation in work)
d = plan.descriptors[plan.offsets[75]:plan.offsets[76]]
def same():
    for f in range(150):
        lib.mvb_composite_stack(frames[f].data_ptr(), W, H, frames.stride(1), bg.ctypes.data, len(d), d.ctypes.data, 0, torch.cuda.current_stream().cuda_stream)
print("frame 75 x150: %.1f us/frame device, %.1f host" % timed(same, 2))

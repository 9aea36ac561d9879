import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
import movis_b200 as mv
from movis_b200 import scenes
from movis_b200.contrib.segmentation import ChromaKey
from movis_b200.render import plan_range, launch_plan
for name, comp in (("s1", scenes.build_s1(mv)), ("s4", scenes.build_s4(mv, ChromaKey)), ("s2", scenes.build_s2(mv))):
    times = np.arange(150) / 30.0
    H, W = comp.frame_shape()
    frames = torch.empty((150, H, W, 4), dtype=torch.uint8, device="cuda")
    t0 = time.perf_counter(); plan = plan_range(comp, times); t_plan = time.perf_counter() - t0
    t0 = time.perf_counter(); plan = plan_range(comp, times); t_plan = time.perf_counter() - t0
    launch_plan(plan, frames, (0, 0, 0, 255)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); launch_plan(plan, frames, (0, 0, 0, 255)); e1.record(); t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{name}: plan {t_plan/150*1e6:.1f} us/frame host, enqueue {t_enq/150*1e6:.1f} us/frame host, device {e0.elapsed_time(e1)/150*1e3:.1f} us/frame")

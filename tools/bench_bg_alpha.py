import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import movis_b200 as mv
from movis_b200 import scenes
from movis_b200.render import plan_range, launch_plan
comp = scenes.build_s2(mv, div=1)
times = (np.arange(150) % 150) / 30.0
H, W = comp.frame_shape()
frames = torch.empty((150, H, W, 4), dtype=torch.uint8, device="cuda")
plan = plan_range(comp, times)
for bg in [(0,0,0,255), (0,0,0,0), (10,20,30,128)]:
    launch_plan(plan, frames, bg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch_plan(plan, frames, bg); e1.record(); torch.cuda.synchronize()
    print(bg, "%.1f us/frame" % (e0.elapsed_time(e1) / 150 * 1e3))

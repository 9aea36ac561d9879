"""Host-side profile of render_frames over many repetitions (finer resolution than host_profile.py)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv
from movis_b200 import scenes
from movis_b200.contrib.segmentation import ChromaKey
which = sys.argv[1] if len(sys.argv) > 1 else "s1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 150
comp = {"s1": lambda: scenes.build_s1(mv), "s2": lambda: scenes.build_s2(mv), "s3": lambda: scenes.build_s3(mv),
        "s4": lambda: scenes.build_s4(mv, ChromaKey), "s5": lambda: scenes.build_s5(mv, ChromaKey)}[which]()
times = np.arange(n) / 30.0
out = comp.render_frames(times)
out = comp.render_frames(times, out=out)
torch.cuda.synchronize()
reps = 20
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(reps):
    out = comp.render_frames(times, out=out)
pr.disable()
host = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"{which}: host {host / reps / n * 1e6:.1f} us/frame (profiled), total {(time.perf_counter() - t0) / reps / n * 1e6:.1f} us/frame")
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

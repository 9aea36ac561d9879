"""Delivered frames/s of stream_frames on S2 (or s5), with the time to the first frame."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv
from movis_b200 import scenes
from movis_b200.render import stream_frames
from movis_b200.contrib.segmentation import ChromaKey
which = sys.argv[1] if len(sys.argv) > 1 else "s2"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
comp = scenes.build_s2(mv) if which == "s2" else scenes.build_s5(mv, ChromaKey)
n = 300
times = np.arange(n) / 30.0 % comp.duration
for rep in range(3):
    t0 = time.perf_counter(); first = None; k = 0
    for f in stream_frames(comp, times, bg_color=(0, 0, 0, 255), batch=batch, copy=False):
        if first is None: first = time.perf_counter() - t0
        k += 1
    dt = time.perf_counter() - t0
    print(f"{which} batch {batch}: {k / dt:.0f} frames/s, {k * 33177600 / dt / 1e9:.1f} GB/s, first frame after {first * 1e3:.1f} ms")

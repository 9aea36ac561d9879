#!/usr/bin/env python
"""Benchmark of the per-frame compositing path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], scene S2 of SURVEY.md 8d): a 3840x2160 composition of 16
synthetic 1920x1080 RGBA layers cycling the blend modes, rotating / scaling, with animated opacity.
One STEP = one pass of the hot path over one batch of synthetic input = rendering the whole 5 s @
30 fps timeline (150 frames): batched keyframe evaluation on the host + one fused kernel launch per
frame.  Metric: composited frames/sec (whole job, all GPUs).

  value     frames/s with sources resident in HBM and frames left in HBM (device timed, CUDA events)
  e2e       frames/s through the public API with HOST buffers: every step uploads the 16 sources
            from pinned host memory and delivers every finished uint8 frame to pinned host memory
  roofline  achieved algorithmic GB/s of the fused kernel k_stack_ws (B_frame = 4*W*H + sum 4*w_l*h_l =
            165 888 000 B, each source texel and each output pixel counted once) / measured HBM peak;
            `traffic` = dram bytes per launch from the committed ncu capture (profiles/roofline_traffic.json).
            One launch = one frame.  mvb_composite_frames alternates the frames of a batch between two
            side streams (forked from / joined into the caller's stream) so that one frame's CTAs fill
            the SMs the previous frame's last CTAs leave idle; `avg_launch_us` is therefore elapsed
            time / launches over back-to-back batches (CUDA events on the caller's stream around the
            fork and the join), and `serial_launch_us` is the same kernel launched one frame at a
            time on the caller's stream (mvb_composite_stack), for comparison with ncu's durations
  cpu_baseline  the CPU oracle (a C + OpenMP port of the reference algorithm) on the host cores

Multi-GPU: launched with torchrun, one rank per GPU; each rank renders its own contiguous 150-frame
range of the timeline (weak scaling); ranks exchange no pixels -- torch.distributed is used only
for the barrier and the max-over-ranks reduction of the elapsed time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "composited_frames_per_sec_4k_16_rgba_layers"
UNIT = "frames/s"
FPS = 30.0
DURATION = 5.0
FRAMES_PER_STEP = 150
WORKLOAD = "S2: 3840x2160, 16x(1920x1080 RGBA) layers, 18 blend modes cycling, animated opacity/scale/rotation"


def measured_hbm_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self) -> None:
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def dist_setup(n_gpus: int):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        from movis_b200.device import bind_host_to_gpu
        bind_host_to_gpu(local)   # pinned frame buffers on the GPU's own NUMA node
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier(world: int) -> None:
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(seconds: float, world: int) -> float:
    import torch
    if world == 1:
        return seconds
    import torch.distributed as dist
    t = torch.tensor([seconds], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def pinned_sources():
    """The 16 S2 source images as numpy arrays backed by pinned host memory."""
    import torch
    from movis_b200 import scenes
    out = []
    for img in scenes.s2_sources(1):
        t = torch.empty(img.shape, dtype=torch.uint8, pin_memory=True)
        t.numpy()[...] = img
        out.append(t.numpy())
    return out


def rank_times(rank: int) -> np.ndarray:
    """Rank r renders the contiguous frame range [150 r, 150 (r + 1)) of the 30 fps timeline,
    wrapped into the 5 s scene (weak scaling: per-GPU work is fixed)."""
    k = np.arange(rank * FRAMES_PER_STEP, (rank + 1) * FRAMES_PER_STEP)
    return (k % int(DURATION * FPS)) / FPS


def run_product(args) -> dict:
    import torch
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.render import launch_plan, plan_range, stream_frames

    rank, local, world = dist_setup(args.gpus)
    sources = pinned_sources()
    comp = scenes.build_s2(mv, div=1, variant=0, sources=sources)
    times = rank_times(rank)
    H, W = comp.frame_shape()
    bg = (0, 0, 0, 255)
    frames = torch.empty((FRAMES_PER_STEP, H, W, 4), dtype=torch.uint8, device="cuda")

    # Two device copies of the sources, alternated frame by frame, so no frame finds the previous
    # frame's texels in L2 (inputs per frame 133 MB + 33 MB out > 126 MB L2 anyway).
    plan0 = plan_range(comp, times)
    base = {int(item.layer.render_device(0.0).data_ptr()): i for i, item in enumerate(comp.layers)}
    twins = [item.layer.render_device(0.0).clone() for item in comp.layers]

    def alternate(plan):
        odd = np.zeros(len(plan.descriptors), dtype=bool)
        for f in range(1, plan.n_frames, 2):
            odd[plan.offsets[f]:plan.offsets[f + 1]] = True
        src = plan.descriptors["src"]
        for p, i in base.items():
            src[odd & (src == p)] = twins[i].data_ptr()
        return plan

    CHUNK = 30   # frames planned per host call, as Composition.render_frames does it (even: keeps the twin parity)

    def step():
        # host: batched keyframes + affine set-up of a chunk; device: the fused launches of the chunk
        # before it, so only the first chunk after an idle device is not overlapped
        for i in range(0, FRAMES_PER_STEP, CHUNK):
            plan = alternate(plan_range(comp, times[i:i + CHUNK]))
            launch_plan(plan, frames[i:i + CHUNK], bg)
        return plan

    for _ in range(max(args.warmup, 3)):
        keep = step()
    barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    ev0.record()
    for _ in range(args.steps):
        keep = step()
    ev1.record()
    barrier(world)
    elapsed = max_over_ranks(ev0.elapsed_time(ev1) / 1e3, world)
    clocks = sampler.stop() if rank == 0 else {}
    total_frames = world * args.steps * FRAMES_PER_STEP
    value = total_frames / elapsed
    launches = args.steps * FRAMES_PER_STEP * 2   # per frame: k_stack_tables (cv2 tables) + k_stack_ws

    # --- roofline of the fused kernel: launches only, plan prebuilt, CUDA events on the launch stream
    plan = alternate(plan_range(comp, times))
    launch_plan(plan, frames, bg)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    k0.record()
    for _ in range(reps):
        launch_plan(plan, frames, bg)
    k1.record()
    torch.cuda.synchronize()
    launch_s = k0.elapsed_time(k1) / 1e3 / (reps * FRAMES_PER_STEP)
    # the same launches one at a time on the current stream (no overlap between frames)
    from movis_b200 import _native
    lib = _native.lib()
    bg_arr = np.asarray(bg, dtype=np.uint8)
    stream = torch.cuda.current_stream().cuda_stream

    def serial():
        for f in range(FRAMES_PER_STEP):
            d = plan.descriptors[plan.offsets[f]:plan.offsets[f + 1]]
            _native.check(lib.mvb_composite_stack(frames[f].data_ptr(), W, H, frames.stride(1), bg_arr.ctypes.data,
                                                  len(d), d.ctypes.data, 0, stream))
    serial()
    torch.cuda.synchronize()
    k0.record()
    serial()
    k1.record()
    torch.cuda.synchronize()
    serial_s = k0.elapsed_time(k1) / 1e3 / FRAMES_PER_STEP
    b_frame = scenes.s2_bytes_per_frame(1)
    peak, peak_src = measured_hbm_peak()
    achieved = b_frame / launch_s / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(), "kernel": "k_stack_ws",
                "algorithmic_bytes_per_launch": b_frame, "avg_launch_us": launch_s * 1e6,
                "serial_launch_us": serial_s * 1e6, "peak_source": peak_src}

    # --- e2e: host sources -> device, render, every frame back to pinned host memory ---------------
    h2d = sum(int(s.nbytes) for s in sources)
    d2h = FRAMES_PER_STEP * H * W * 4
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        comp.cache.clear()
        for item in comp.layers:
            item.layer._device_image = None          # forces the H2D upload of every source
        n = 0
        for _frame in stream_frames(comp, times, bg_color=bg, batch=10, copy=False):
            n += 1
        return n

    e2e_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert e2e_step() == FRAMES_PER_STEP
    barrier(world)
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0, world)
    e2e = {"value": world * e2e_steps * FRAMES_PER_STEP / e2e_elapsed, "unit": UNIT,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "api": "movis_b200.render.stream_frames (Composition._write_video loop)"}

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "fps": FPS,
                   "l2": "inputs larger than L2 (133 MB sources + 33 MB frame per launch, sources double-buffered)",
                   "parallelism": f"frame-range sharding x{world}, no collective"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if rank == 0 and world == 1:
        result["cpu_baseline"] = cpu_baseline(sample_frames=6)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return result if rank == 0 else {}


def cpu_baseline(sample_frames: int) -> dict:
    """The oracle port of the reference algorithm on the host cores, on every 25th frame of the
    same timeline (bounded sample), keyframes evaluated per frame like the reference does."""
    import movis_b200 as mv
    from movis_b200 import scenes
    from oracle import oracle as O
    from oracle import scene_eval
    # torchrun exports OMP_NUM_THREADS=1; the baseline is meant to use every host thread we may run on
    O.set_num_threads(len(os.sched_getaffinity(0)))
    comp = scenes.build_s2(mv, div=1, variant=0)
    times = (np.arange(sample_frames) * 25 % 150) / FPS
    scene_eval.render(comp, float(times[0]), (0, 0, 0, 255))  # warm up (page in sources)
    t0 = time.perf_counter()
    for t in times:
        scene_eval.render(comp, float(t), (0, 0, 0, 255))
    dt = time.perf_counter() - t0
    return {"value": sample_frames / dt, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": f"{sample_frames} frames of the S2 timeline (every 25th), C oracle with OpenMP on all host threads"}


def run_reference(args) -> dict:
    """--impl reference: the reference's CPU algorithm (oracle port; the Python reference cannot
    travel to the GPU box) on all host threads.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    per_step = 2
    base = None
    for _ in range(max(args.warmup, 1)):
        base = cpu_baseline(per_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        base = cpu_baseline(per_step)
    dt = time.perf_counter() - t0
    # cpu_baseline() renders per_step + 1 frames (one warm-up frame) per call
    value = args.steps * (per_step + 1) / dt
    cb = dict(base)
    cb["value"] = value
    cb["sample"] = f"{per_step + 1} frames of the S2 timeline per step; " + cb["sample"]
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step + 1, "fps": FPS},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    result = run_reference(args) if args.impl == "reference" else run_product(args)
    if result:
        print(json.dumps(result))


if __name__ == "__main__":
    main()

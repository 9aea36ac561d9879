#!/usr/bin/env python
"""Benchmark of the per-frame compositing path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--quick]

Headline workload (BASELINE.json configs[1], scene S2 of SURVEY.md 8d): a 3840x2160 composition of 16
synthetic 1920x1080 RGBA layers cycling the blend modes, rotating / scaling, with animated opacity.
One STEP = one pass of the hot path over one batch of synthetic input = rendering 150 frames through
the public batched renderer (movis_b200.render.render_frames: batched keyframe evaluation on the host,
one persistent fused kernel launch per run of up to 32 frames).  Metric: composited frames/sec (whole
job, all GPUs).

  value     frames/s with sources resident in HBM and frames left in HBM (device timed, CUDA events)
  e2e       frames/s through the public API with HOST buffers: every step uploads the 16 sources
            from pinned host memory and delivers every finished uint8 frame to pinned host memory
            (movis_b200.render.stream_frames, the loop of Composition._write_video)
  roofline  achieved algorithmic GB/s of the fused kernel k_stack_frames / measured HBM peak.  One launch
            = one run of `frames_per_launch` frames; algorithmic bytes per frame B_frame = 4*W*H +
            sum 4*w_l*h_l = 165 888 000 B (each source texel and each output pixel counted once).  Timed
            live with CUDA events on the launch stream over prebuilt plans (the events also bracket each
            run's 100 KB record upload and its preparation kernel k_stack_prep, < 2 % of the run).
            `traffic` = dram bytes per launch scaled from the committed ncu capture
            (profiles/roofline_traffic.json; `traffic_source` says so: it is not measured by this run).
  configs   device frames/s and roofline fraction of the other BASELINE.json configs (S1, S2d, S3, S4 at
            draft levels 1/2/4, S5), each from a few hundred ms of rendering
  operators_4k  the stand-alone effect kernels at 4K through the C ABI (us per call, GB/s, roofline fraction)
  s5        BASELINE.json configs[4]: the 1800-frame 4K timeline, frame-sharded over the ranks (STRONG
            scaling: total work fixed): device and delivered frames/s of the whole job
  cpu_baseline  the reference's algorithm on the host cores (rank 0, N = 1): the real reference when it
            is importable (/root/reference, build container only), else its C + OpenMP oracle port

Multi-GPU: launched with torchrun, one rank per GPU.  The job is ONE fixed timeline of 150 * N frames
(the 5 s scene sampled at 30 * N fps); rank r renders the contiguous index range
render.shard_range(150 N, r, N) of it, so ranks render DIFFERENT frames and per-GPU work is fixed
(weak scaling).  Ranks exchange no pixels -- torch.distributed is used only for the barrier, the
max-over-ranks reduction of the elapsed time and the all-gather of eight numbers below.
The delivered-frame legs (e2e, s5 delivered) split the same timeline in proportion to each rank's
device -> host copy rate measured with all ranks copying at once (render.d2h_rate, e2e.shard_weights_GBps):
on this pool's 8-GPU boxes half of the GPUs get half the PCIe rate of the others under load.
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
BG = (0, 0, 0, 255)


def bench_config(world: int) -> dict:
    return {"workload": WORKLOAD, "frames_per_step": FRAMES_PER_STEP, "fps": FPS,
            "l2": "inputs larger than L2 (133 MB sources + 33 MB frame per frame; the sources are "
                  "double-buffered and alternated frame by frame, so no frame finds the previous frame's texels in L2)",
            "parallelism": f"frame-range sharding x{world} of one timeline (render.shard_range), no collective"}


def measured_hbm_peak() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def recorded_traffic_per_frame():
    """dram bytes per FRAME of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            d = json.load(fh)
            return d.get("dram_bytes_per_frame"), d.get("source")
    except Exception:
        return None, None


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


def dist_setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        from movis_b200.device import bind_host_to_gpu
        torch.cuda.set_device(local)
        dist_setup.binding = bind_host_to_gpu(local)   # pinned frame buffers on the GPU's own NUMA node
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


def gather_floats(values, world: int):
    """Every rank's list of floats, as a list of lists (diagnostics: per-rank step times and frame counts)."""
    import torch
    if world == 1:
        return [list(values)]
    import torch.distributed as dist
    mine = torch.tensor(values, dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [[round(float(v), 3) for v in t.tolist()] for t in out]


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


def timeline(world: int, rank: int, weights=None, comp=None) -> np.ndarray:
    """This rank's share of the job: ONE timeline of 150 * world frames (the 5 s scene at 30 * world
    fps), contiguous index range render.shard_range(150 world, rank, world[, weights][, costs]).  With
    `comp` the cuts balance the frames' estimated cost (render.frame_costs: the layers of S2 grow by 44 %
    over the timeline) instead of their count; every rank computes the same cuts from the same scene."""
    from movis_b200.render import frame_costs, shard_range
    total = FRAMES_PER_STEP * world
    all_times = np.arange(total, dtype=np.float64) / (FPS * world)
    costs = frame_costs(comp, all_times) if (comp is not None and world > 1) else None
    mine = shard_range(total, rank, world, weights, costs)
    return all_times[mine.start:mine.stop]


def delivery_weights(world: int):
    """Per-rank device -> host copy rates with every rank copying at once (GB/s), the weights of the
    delivered-frame legs: on the boxes of this pool half of the GPUs get half the PCIe rate of the others
    when all eight copy (profiles/r02_pcie_concurrent.json), and equal shares would wait for them."""
    if world == 1:
        return None
    import torch
    import torch.distributed as dist
    from movis_b200.render import d2h_rate
    barrier(world)
    mine = torch.tensor([d2h_rate() / 1e9], dtype=torch.float64, device="cuda")
    rates = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rates, mine)
    return [round(float(r.item()), 2) for r in rates]


# ---------------------------------------------------------------------------------------------
# the other BASELINE.json configs (device frames/s + roofline fraction, a few hundred ms each)
# ---------------------------------------------------------------------------------------------

def s3_bytes_per_frame() -> int:
    # steady state: the three effect outputs are static images (cached once per range), so per frame the
    # inner frame reads its colour background + them and is written; the parent reads it + its own
    # background and is written (SURVEY 8d without the once-per-range effect stages)
    inner = 4 * 1920 * 1080
    outs = 4 * (482 * 482 + 510 * 510 + 482 * 482)
    outer = 4 * 3840 * 2160
    return (inner + outs + inner) + (outer + inner + outer)


def s4_bytes_per_frame(level: int) -> int:
    # frame + keyed 4K frame + the bg region sampled, except that a draft frame cannot need more
    # of a source than the four 4-byte taps of each of its pixels: at level 4 the samples are 4 texels apart
    # and three quarters of the source rows are never touched, so counting whole sources would overstate the
    # traffic (and put the "roofline fraction" above 1)
    out_px = (3840 // level) * (2160 // level)
    return 4 * out_px + 2 * min(33177600, 16 * out_px)


def s5_bytes_per_frame() -> int:
    from movis_b200 import scenes
    inner = 4 * 1920 * 1080
    outs = 4 * (482 * 482 + 510 * 510 + 482 * 482)
    return scenes.s2_bytes_per_frame(1) + (inner + outs + inner) + inner + 33177600 + 6 * inner


def time_scene(comp, times: np.ndarray, level: int = 1, reps: int = 2) -> float:
    """Seconds per frame of `comp` over `times` through Composition.render_frames (second pass onwards)."""
    import torch
    with comp.preview(level=level):
        out = comp.render_frames(times, bg_color=BG)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = comp.render_frames(times, bg_color=BG, out=out)
        e1.record()
        torch.cuda.synchronize()
    del out
    return e0.elapsed_time(e1) / 1e3 / (reps * len(times))


def other_configs(peak: float) -> dict:
    import torch
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.contrib.segmentation import ChromaKey
    full = np.arange(0, DURATION, 1 / FPS)
    out = {}

    def record(name, spf, nbytes, frames, note=""):
        out[name] = {"device_frames_per_s": 1.0 / spf, "us_per_frame": spf * 1e6, "algorithmic_bytes_per_frame": nbytes,
                     "achieved_GBps": nbytes / spf / 1e9, "roofline_frac": nbytes / spf / 1e9 / peak, "frames": frames}
        if note:
            out[name]["note"] = note

    comp = scenes.build_s1(mv)
    record("S1 (configs[0]: 1080p, bg + 8 rotating layers)", time_scene(comp, full, reps=4), scenes.s1_bytes_per_frame(), len(full))
    comp = scenes.build_s2d(mv)
    record("S2d (dense: 17 full-frame 4K layers)", time_scene(comp, full[:60]), scenes.s2d_bytes_per_frame(), 60)
    del comp
    torch.cuda.empty_cache()
    comp = scenes.build_s3(mv)
    record("S3 (configs[2]: nested 1080p comp with blur / shadow / glow in a 4K parent)", time_scene(comp, full),
           s3_bytes_per_frame(), len(full), "effect outputs are static: computed once per range")
    comp = scenes.build_s4(mv, ChromaKey)
    for level in (1, 2, 4):
        record(f"S4 L={level} (configs[3]: chroma-keyed 4K sequence over a 4.2K background)", time_scene(comp, full, level=level),
               s4_bytes_per_frame(level), len(full), "the chroma key runs as its own pass once per distinct sequence frame (8 of them; cached through ChromaKey.get_key)")
    del comp
    torch.cuda.empty_cache()
    return out


def operator_timings(peak: float) -> dict:
    """The stand-alone effect operators at 4K through the C ABI with preallocated buffers (CUDA events over
    40 back-to-back calls on four alternating inputs): microseconds per call and algorithmic GB/s
    (input + output once; blur counts as one stage)."""
    import torch
    from movis_b200 import _native
    L = _native.lib()
    H, W, reps = 2160, 3840, 40
    g = torch.Generator(device="cuda").manual_seed(0)
    src = [torch.randint(0, 256, (H, W, 4), dtype=torch.uint8, device="cuda", generator=g) for _ in range(4)]
    for t in src:
        t[: H // 2, :, 0] = 0
        t[: H // 2, :, 1] = 255
        t[: H // 2, :, 2] = 0
    dst = torch.empty((H + 200, W + 200, 4), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    out = {}

    def timed(name, fn, nbytes):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        out[name] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1), "roofline_frac": round(nbytes / us / 1e3 / peak, 3)}

    lo, hi = np.array([50, 178, 178], np.uint8), np.array([70, 255, 255], np.uint8)
    timed("ChromaKey()", lambda i: _native.check(L.mvb_effect_chroma_key_rgba8(
        src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), lo.ctypes.data, hi.ctypes.data, stream)), 2 * H * W * 4)
    col = np.array([255, 0, 0], np.uint8)
    timed("FillColor", lambda i: _native.check(L.mvb_effect_fill_color_rgba8(
        src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), col.ctypes.data, stream)), 2 * H * W * 4)
    x = np.arange(256, dtype=np.uint8)
    lut = np.ascontiguousarray((x.astype(np.int32) * 3 // 4).astype(np.uint8))
    timed("HSLShift", lambda i: _native.check(L.mvb_effect_hsl_shift_rgba8(
        src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), lut.ctypes.data, lut.ctypes.data, lut.ctypes.data,
        stream)), 2 * H * W * 4)
    for radius in (3.0, 10.0):
        k = _native.blur_ksize(radius)
        wts = _native.gaussian_weights_q8(radius, k)
        scratch = torch.empty(L.mvb_effect_blur_scratch_bytes(W, H, k, 4), dtype=torch.uint8, device="cuda")
        glow = np.clip(1.3 * x.astype(np.float32), 0, 255).astype(np.uint8)
        nbytes = H * W * 4 + (H + 2 * k) * (W + 2 * k) * 4
        timed(f"GaussianBlur({radius:g})", lambda i: _native.check(L.mvb_effect_blur_rgba8(
            src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), k, wts.ctypes.data, _native.MVB_BLUR_PLAIN,
            None, scratch.data_ptr(), stream)), nbytes)
        timed(f"Glow({radius:g}, 1.3)", lambda i: _native.check(L.mvb_effect_blur_rgba8(
            src[i & 3].data_ptr(), W, H, W * 4, dst.data_ptr(), dst.stride(0), k, wts.ctypes.data, _native.MVB_BLUR_GLOW,
            glow.ctypes.data, scratch.data_ptr(), stream)), nbytes)
    del src, dst
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------

def run_product(args) -> dict:
    import torch
    import movis_b200 as mv
    from movis_b200 import _native, scenes
    from movis_b200.render import launch_plan, plan_range, render_frames, shard_range, stream_frames

    rank, local, world = dist_setup()
    lib = _native.lib()
    sources = pinned_sources()
    comp = scenes.build_s2(mv, div=1, variant=0, sources=sources)
    times = timeline(world, rank, comp=comp)
    H, W = comp.frame_shape()
    frames = torch.empty((len(times), H, W, 4), dtype=torch.uint8, device="cuda")

    # Two device copies of the sources, alternated frame by frame through render_frames' plan hook, so no
    # frame finds the previous frame's texels in L2 (inputs per frame 133 MB + 33 MB out > 126 MB L2 anyway).
    base = {int(item.layer.render_device(0.0).data_ptr()): i for i, item in enumerate(comp.layers)}
    twins = [item.layer.render_device(0.0).clone() for item in comp.layers]

    def alternate(plan, first_index):
        odd = np.zeros(len(plan.descriptors), dtype=bool)
        for f in range(plan.n_frames):
            if (first_index + f) & 1:
                odd[plan.offsets[f]:plan.offsets[f + 1]] = True
        src = plan.descriptors["src"]
        for p, i in base.items():
            src[odd & (src == p)] = twins[i].data_ptr()
        return plan

    def step():
        # the public batched path: host plans a chunk (batched keyframes + affine set-up) while the device
        # renders the one before it
        render_frames(comp, times, bg_color=BG, out=frames, plan_hook=alternate)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier(world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    launches0 = int(lib.mvb_kernel_launches())
    ev0.record()
    host0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_s = time.perf_counter() - host0
    ev1.record()
    barrier(world)
    launches = int(lib.mvb_kernel_launches()) - launches0
    elapsed = max_over_ranks(ev0.elapsed_time(ev1) / 1e3, world)
    per_rank = gather_floats([ev0.elapsed_time(ev1) / args.steps, host_s / args.steps * 1e3, float(len(times))], world)
    clocks = sampler.stop() if rank == 0 else {}
    total_frames = world * args.steps * FRAMES_PER_STEP
    value = total_frames / elapsed

    # --- roofline of the fused kernel: launches only, plans prebuilt, CUDA events on the launch stream
    from movis_b200.render import PLAN_CHUNK
    n_mine = len(times)   # this rank's frames per step (150 at N = 1; cost-balanced around 150 otherwise)
    plans = [alternate(plan_range(comp, times[i:i + PLAN_CHUNK]), i) for i in range(0, n_mine, PLAN_CHUNK)]

    def launch_all():
        for k, p in enumerate(plans):
            launch_plan(p, frames[k * PLAN_CHUNK:k * PLAN_CHUNK + p.n_frames], BG)

    launch_all()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    n0 = int(lib.mvb_kernel_launches())
    k0.record()
    for _ in range(reps):
        launch_all()
    k1.record()
    torch.cuda.synchronize()
    runs = (int(lib.mvb_kernel_launches()) - n0) // 2 // reps      # each run = k_stack_prep + k_stack_frames
    frame_s = k0.elapsed_time(k1) / 1e3 / (reps * n_mine)
    # the same frames one per call (mvb_composite_stack: a run of one frame each), for comparison with
    # per-frame launch schemes
    bg_arr = np.asarray(BG, dtype=np.uint8)
    stream = torch.cuda.current_stream().cuda_stream
    plan = plans[0]

    def serial():
        for f in range(plan.n_frames):
            d = plan.descriptors[plan.offsets[f]:plan.offsets[f + 1]]
            _native.check(lib.mvb_composite_stack(frames[f].data_ptr(), W, H, frames.stride(1), bg_arr.ctypes.data,
                                                  len(d), d.ctypes.data, 0, stream))
    serial()
    torch.cuda.synchronize()
    k0.record()
    serial()
    k1.record()
    torch.cuda.synchronize()
    serial_s = k0.elapsed_time(k1) / 1e3 / plan.n_frames
    b_frame = scenes.s2_bytes_per_frame(1)
    peak, peak_src = measured_hbm_peak()
    achieved = b_frame / frame_s / 1e9
    frames_per_launch = n_mine / max(runs, 1)
    traffic_frame, traffic_src = recorded_traffic_per_frame()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None if traffic_frame is None else traffic_frame * frames_per_launch,
                "traffic_source": None if traffic_frame is None else f"committed ncu capture, not measured by this run ({traffic_src})",
                "kernel": "k_stack_frames", "frames_per_launch": frames_per_launch,
                "algorithmic_bytes_per_launch": b_frame * frames_per_launch, "algorithmic_bytes_per_frame": b_frame,
                "avg_launch_us": frame_s * frames_per_launch * 1e6, "us_per_frame": frame_s * 1e6,
                "single_frame_call_us": serial_s * 1e6, "peak_source": peak_src}

    # --- e2e: host sources -> device, render, every frame back to pinned host memory ---------------
    h2d = sum(int(s.nbytes) for s in sources)
    d2h = FRAMES_PER_STEP * H * W * 4
    e2e_steps = max(1, min(args.steps, 3))

    weights = delivery_weights(world)
    times_e2e = timeline(world, rank, weights)     # this rank's share of the SAME 150 * world frames

    def e2e_step():
        comp.cache.clear()
        for item in comp.layers:
            item.layer._device_image = None          # forces the H2D upload of every source
        n = 0
        for _frame in stream_frames(comp, times_e2e, bg_color=BG, copy=False):
            n += 1
        return n

    e2e_step()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert e2e_step() == len(times_e2e)
    barrier(world)
    e2e_elapsed = max_over_ranks(time.perf_counter() - t0, world)
    e2e = {"value": world * e2e_steps * FRAMES_PER_STEP / e2e_elapsed, "unit": UNIT,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "d2h_GBps": world * e2e_steps * d2h / e2e_elapsed / 1e9,
           "api": "movis_b200.render.stream_frames (Composition._write_video loop)",
           "shard_weights_GBps": weights}

    result = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(world),
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "host_binding": getattr(dist_setup, "binding", None),
        "per_rank": {"device_ms_per_step": [r[0] for r in per_rank], "host_ms_per_step": [r[1] for r in per_rank],
                     "frames_per_step": [int(r[2]) for r in per_rank]},
    }

    if not args.quick:
        # --- BASELINE.json configs[4]: the 1800-frame timeline, frame-sharded (strong scaling) ---------
        frames = None
        twins.clear()
        comp = None
        torch.cuda.empty_cache()
        result["s5"] = run_s5(world, rank, weights)
        if rank == 0 and world == 1:
            result["configs"] = other_configs(peak)
            result["operators_4k"] = operator_timings(peak)
            result["cpu_baseline"] = cpu_baseline()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return result if rank == 0 else {}


def run_s5(world: int, rank: int, weights=None) -> dict:
    import torch
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.contrib.segmentation import ChromaKey
    from movis_b200.render import render_frames, shard_range, stream_frames
    N = 1800
    comp = scenes.build_s5(mv, ChromaKey)
    from movis_b200.render import frame_costs
    all_times = np.arange(N, dtype=np.float64) / FPS
    # device leg: contiguous ranges of equal estimated cost (the layers grow over the timeline)
    mine = shard_range(N, rank, world, costs=frame_costs(comp, all_times) if world > 1 else None)
    times = all_times[mine.start:mine.stop]
    for _ in stream_frames(comp, times[:16], bg_color=BG, copy=False):   # warm-up: uploads, effect caches
        pass
    CALL = 240   # frames per render_frames call (8 GB of frames resident; the call plans 240 frames at once)
    out = torch.empty((CALL,) + comp.frame_shape() + (4,), dtype=torch.uint8, device="cuda")
    # warm-up: one untimed pass over this rank's whole range at the call size of the timed loop (uploads of the
    # 8 sequence frames and their keyed versions, allocator blocks of the nested composition's batches), so
    # that the timed pass runs with everything resident, as SURVEY.md 8d specifies for S4 / S5
    for i in range(0, len(times), CALL):
        chunk = times[i:i + CALL]
        render_frames(comp, chunk, bg_color=BG, out=out[:len(chunk)])
    barrier(world)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(0, len(times), CALL):
        chunk = times[i:i + CALL]
        render_frames(comp, chunk, bg_color=BG, out=out[:len(chunk)])
    ev1.record()
    barrier(world)
    device_s = max_over_ranks(ev0.elapsed_time(ev1) / 1e3, world)
    del out
    times_d = np.asarray(shard_range(N, rank, world, weights), dtype=np.float64) / FPS   # delivered leg: weighted ranges
    barrier(world)
    t0 = time.perf_counter()
    n = 0
    for _ in stream_frames(comp, times_d, bg_color=BG, copy=False):
        n += 1
    barrier(world)
    delivered_s = max_over_ranks(time.perf_counter() - t0, world)
    peak, _ = measured_hbm_peak()
    b = s5_bytes_per_frame()
    return {"workload": "S5 (configs[4]): 60 s 4K 30 fps, 1800 frames, 24 layers + nested effects + keyed sequence",
            "scaling": "strong", "frames": N, "frames_per_rank": len(mine),
            "device_frames_per_s": N / device_s, "delivered_frames_per_s": N / delivered_s,
            "algorithmic_bytes_per_frame": b, "roofline_frac_per_gpu": b * N / device_s / 1e9 / peak / world}


# ---------------------------------------------------------------------------------------------
# the reference's CPU implementation of the path
# ---------------------------------------------------------------------------------------------

SAMPLE_TIMES = (0.5, 2.5, 4.5)   # the bounded sample of the S2 timeline the CPU arms render per step


class CpuArm:
    """The reference's own per-frame render of S2 on the host cores: the unmodified reference when it is
    importable (oracle.ref_harness, build container only -- /root/reference does not travel to the GPU
    box), else the C + OpenMP oracle port of the same algorithm on all host threads.  The scene is built
    ONCE; every step renders the same time samples."""

    def __init__(self) -> None:
        import movis_b200 as mv
        from movis_b200 import scenes
        from oracle import oracle as O
        from oracle import ref_harness, scene_eval
        self.kind, self.cores = "port", len(os.sched_getaffinity(0))
        if ref_harness.available():
            ref = ref_harness.import_reference()
            self.kind = "reference"
            comp = scenes.build_s2(ref, div=1, variant=0)
            self.render = lambda t: comp(float(t), bg_color=BG)
            try:
                import cv2
                self.cores = cv2.getNumThreads()   # numpy and PIL are single-threaded; cv2 uses its pool
            except Exception:
                self.cores = 1
            self.how = "unmodified reference (numpy / cv2 / PIL), one process, cv2 default threads"
        else:
            # torchrun exports OMP_NUM_THREADS=1; the baseline is meant to use every host thread we may run on
            O.set_num_threads(self.cores)
            self.cores = O.num_threads()
            comp = scenes.build_s2(mv, div=1, variant=0)
            self.render = lambda t: scene_eval.render(comp, float(t), BG)
            self.how = "C + OpenMP oracle port of the reference algorithm on all host threads"
        self.render(SAMPLE_TIMES[0])   # warm up (page in the sources)

    def step(self) -> float:
        t0 = time.perf_counter()
        for t in SAMPLE_TIMES:
            self.render(t)
        return time.perf_counter() - t0

    def describe(self) -> str:
        return f"{len(SAMPLE_TIMES)} frames of the S2 timeline (t = {', '.join(str(t) for t in SAMPLE_TIMES)} s) per step; {self.how}"


def cpu_baseline() -> dict:
    arm = CpuArm()
    reps = 1 if arm.kind == "reference" else 4
    dt = sum(arm.step() for _ in range(reps))
    return {"value": reps * len(SAMPLE_TIMES) / dt, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.describe()}


def run_reference(args) -> dict:
    """--impl reference: the reference's CPU implementation of the path, same metric and config as the
    product arm; each step renders a bounded sample of the workload (the same frames every step, scene
    built once outside the timed loop)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    arm = CpuArm()
    for _ in range(max(args.warmup, 1) if arm.kind == "port" else 0):
        arm.step()
    steps = args.steps if arm.kind == "port" else min(args.steps, 2)   # the real reference needs ~8 s per 4K frame
    dt = sum(arm.step() for _ in range(steps))
    value = steps * len(SAMPLE_TIMES) / dt
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(args.warmup, 1) if arm.kind == "port" else 0, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": bench_config(int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": arm.describe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline numbers only (skip S5, the other configs and the CPU baseline)")
    args = ap.parse_args()
    result = run_reference(args) if args.impl == "reference" else run_product(args)
    if result:
        print(json.dumps(result))


if __name__ == "__main__":
    main()

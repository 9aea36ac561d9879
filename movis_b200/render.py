"""Frame-range renderer: the GPU replacement for the loop of ``Composition._write_video``
(reference: composition.py:405-413) and the unit that is sharded across GPUs.

Frames are independent in time, so a range is rendered as a batch:
  1. keyframes of every layer are evaluated for the whole range at once (``Transform.sample``),
     the affine set-up is one stacked numpy computation per layer (bit-identical to the per-frame
     path), and the result is one flat table of ``mvb_layer`` descriptors;
  2. one C call (``mvb_composite_frames``) enqueues the fused kernel for every frame of the batch;
  3. ``stream_frames`` moves finished frames to pinned host memory on a copy stream while the
     next batch renders, and yields them in order.

Multi-GPU: ``shard_range`` gives rank r of G the contiguous slice of frame indices it owns; ranks
never exchange pixels (no collective on the data path).
"""
from __future__ import annotations

from typing import Iterator, Sequence

import numpy as np
import torch

from . import _native
from .device import require_cuda, stream_ptr
from .layer.composition import Composition, LayerItem, fill_descriptors, plan_affine, plan_affine_xy


PLAN_CHUNK = 256                 # most frames planned per host step of render_frames
NESTED_BATCH_BYTES = 2 << 30     # budget for the frames of a nested composition rendered per step


def plan_chunk(comp: Composition) -> int:
    """Frames planned (and launched) per host step.  Keyframe evaluation is a handful of numpy calls per
    layer whatever the number of frames, so long steps keep the host ahead of the device (S1: 37 us per
    frame of host work with 64-frame steps against 16 us of device work); a nested composition's frames
    for a whole step are resident at once, which bounds the step by memory instead."""
    worst = 0
    for item in comp.layers:
        if isinstance(item.layer, Composition):
            h, w = item.layer.frame_shape()
            worst = max(worst, h * w * 4)
    if worst == 0:
        return PLAN_CHUNK
    return int(min(PLAN_CHUNK, max(16, NESTED_BATCH_BYTES // worst)))


def shard_range(n_frames: int, rank: int, world_size: int, weights: Sequence[float] | None = None,
                costs: Sequence[float] | None = None) -> range:
    """Contiguous frame indices owned by ``rank``: [floor(r N / G), floor((r + 1) N / G)).

    With ``weights`` (one positive number per rank, the same list on every rank) the cut points follow the
    cumulative weights instead, so a rank that delivers frames faster owns a longer -- still contiguous --
    range.  On the 8-GPU boxes measured here the device -> host copy rate differs 2 : 1 between GPUs when
    all of them copy at once (profiles/r02_pcie_concurrent.json), and the delivered frame rate is the
    slowest rank's; ``d2h_rate`` measures the weights.

    With ``costs`` (one non-negative number per frame, e.g. ``frame_costs``) the cuts balance the summed
    cost instead of the frame count: frames late in S2's timeline carry 40 % more layer pixels than early
    ones, and with equal counts every rank waits for the one that owns the end of the timeline."""
    assert 0 <= rank < world_size
    if weights is None and costs is None:
        return range((rank * n_frames) // world_size, ((rank + 1) * n_frames) // world_size)
    w = np.ones(world_size) if weights is None else np.asarray(weights, dtype=np.float64)
    assert w.shape == (world_size,) and np.all(w > 0), "one positive weight per rank"
    share = np.concatenate([[0.0], np.cumsum(w)]) / w.sum()          # cumulative share of the work per rank
    if costs is None:
        cuts = np.floor(share * n_frames + 1e-9).astype(np.int64)
    else:
        c = np.asarray(costs, dtype=np.float64)
        assert c.shape == (n_frames,) and np.all(c >= 0)
        acc = np.concatenate([[0.0], np.cumsum(c)])
        total = acc[-1] if acc[-1] > 0 else 1.0
        # first frame index whose cumulative cost reaches the rank's share (monotone, so ranges stay contiguous)
        cuts = np.searchsorted(acc / total, share - 1e-12, side="left").astype(np.int64)
        cuts[0] = 0
    cuts[-1] = n_frames
    cuts = np.maximum.accumulate(cuts)
    return range(int(cuts[rank]), int(cuts[rank + 1]))


def frame_costs(comp: Composition, times: Sequence[float]) -> np.ndarray:
    """Relative render cost of every frame of ``times``: frame pixels + the pixels every composited layer
    covers (what the fused kernel walks).  Host only, a few numpy calls per layer for the whole
    range; every rank computes the same numbers from the same scene, so ``shard_range(…, costs=…)`` needs no
    communication."""
    times = np.asarray(times, dtype=np.float64)
    H, W = comp.frame_shape()
    L = comp.preview_level
    cost = np.full(len(times), float(H * W))
    for item in comp.layers:
        if not item.visible:
            continue
        layer_times = times - item.offset
        idx = np.nonzero(~((layer_times < item.start_time) | (item.end_time <= layer_times)))[0]
        if len(idx) == 0:
            continue
        size = getattr(item.layer, "size", None)
        if size is None and isinstance(item.layer, Composition):
            size = item.layer.size
        if size is None:
            continue   # (unknown source size: left out of the estimate)
        plan = plan_affine(item.transform.sample(layer_times[idx]), (int(size[0]) // L or 1, int(size[1]) // L or 1)
                           if isinstance(item.layer, Composition) else (int(size[0]), int(size[1])), L)
        # pixels the layer covers: source area times |det| of the forward map, never more than its bbox
        # (the bbox of a rotated layer is mostly empty tiles, which the kernel culls)
        m = plan.matrix.reshape(-1, 6)
        w_src, h_src = (int(size[0]) // L or 1, int(size[1]) // L or 1) if isinstance(item.layer, Composition) else (int(size[0]), int(size[1]))
        covered = np.abs(m[:, 0] * m[:, 4] - m[:, 1] * m[:, 3]) * float(w_src * h_src)
        bbox = plan.bbox[:, 2].astype(np.float64) * plan.bbox[:, 3]
        cost[idx] += np.where(plan.valid, np.minimum(np.minimum(covered, bbox), 4.0 * H * W), 0.0)
    return cost


def d2h_rate(nbytes: int = 256 << 20, repeats: int = 3) -> float:
    """Bytes per second of a pinned device -> host copy on the current device, as a weight for
    ``shard_range``.  Call it on every rank at the same time (after a barrier): what matters is the rate
    each GPU gets while the others copy too."""
    require_cuda()
    import time
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(repeats):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    return repeats * nbytes / (time.perf_counter() - t0)


class RangePlan:
    """Host-side result of planning N frames: descriptors in frame-major paint order."""

    def __init__(self, descriptors: np.ndarray, offsets: np.ndarray, keepalive: list) -> None:
        self.descriptors = descriptors  # _native.LAYER_DTYPE, len == offsets[-1]
        self.offsets = offsets          # int32 (N + 1,)
        self.keepalive = keepalive      # source tensors referenced by the descriptors

    @property
    def n_frames(self) -> int:
        return len(self.offsets) - 1

    def slice(self, a: int, b: int) -> "RangePlan":
        """Frames [a, b) of this plan (views of the same tables)."""
        lo, hi = int(self.offsets[a]), int(self.offsets[b])
        return RangePlan(self.descriptors[lo:hi], (self.offsets[a:b + 1] - self.offsets[a]).astype(np.int32), self.keepalive)


def _static_content(item: LayerItem) -> bool:
    """True when the post-effect image of ``item`` cannot change while it is visible, so one
    device image serves the whole range."""
    layer = item.layer
    from .layer.media import Image
    if not isinstance(layer, Image):
        return False
    for e in item.effects:
        attrs = getattr(e, "attributes", None)
        if not hasattr(e, "apply_device") or attrs is None:
            return False
        if not all(a.is_static for a in attrs.values()):
            return False
    return True


def _effects_time_invariant(item: LayerItem) -> bool:
    """True when the effect chain of ``item`` maps equal input images to equal output images whatever
    the time: every effect runs on the device, defines ``get_key`` and has no animated attribute."""
    for e in item.effects:
        if not hasattr(e, "apply_device") or not hasattr(e, "get_key"):
            return False
        attrs = getattr(e, "attributes", None)
        if attrs is not None and not all(a.is_static for a in attrs.values()):
            return False
    return True


def plan_range(comp: Composition, times: Sequence[float]) -> RangePlan:
    """Evaluate the scene for every element of ``times`` (all inside ``[0, duration)``)."""
    times = np.asarray(times, dtype=np.float64)
    n = len(times)
    L = comp.preview_level
    per_layer = []  # (frame indices, descriptor rows) per layer, in paint order
    static_jobs = []  # (slot in per_layer, frame indices, inside mask, image, samples) of the static-content layers
    keep: list = []
    for item in comp.layers:
        if not item.visible:
            continue
        layer_times = times - item.offset
        window = ~((layer_times < item.start_time) | (item.end_time <= layer_times))
        idx = np.nonzero(window)[0]
        if len(idx) == 0:
            continue
        samples = item.transform.sample(layer_times[idx])
        if _static_content(item):
            # Image.render_device is None outside [0, image.duration)
            inside = (layer_times[idx] >= 0) & (layer_times[idx] < item.layer.duration)
            image = item.fg_image_device(float(times[idx[inside][0]]), comp.cache) if inside.any() else None
            if image is None:
                continue
            keep.append(image)
            static_jobs.append((len(per_layer), idx, inside, image, samples))
            per_layer.append(None)   # filled in below, by ONE affine plan for all static layers
            continue
        # general case: the image may change per frame (sequences, nested comps, animated effects)
        rows = np.zeros(len(idx), dtype=_native.LAYER_DTYPE)
        ok = np.zeros(len(idx), dtype=bool)
        by_size: dict[tuple[int, int], list[int]] = {}
        images = [None] * len(idx)
        if isinstance(item.layer, Composition):
            # Nested composition (composition.py: an inner Composition is just a layer, rendered with
            # the default transparent background, its own cache and preview level): its frames for
            # the whole range are planned and rendered as ONE batch, recursively, instead of one
            # Python-level render per frame.
            inner = item.layer
            lt = layer_times[idx]
            inside = np.nonzero((lt >= 0.0) & (lt < inner.duration))[0]
            if len(inside):
                batch = render_frames(inner, lt[inside], bg_color=(0, 0, 0, 0))
                keep.append(batch)
                for k, j in enumerate(inside):
                    images[j] = item.apply_effects_device(batch[k], float(lt[j]))
        elif hasattr(item.layer, "get_keys") and _effects_time_invariant(item):
            # one fetch per DISTINCT source image of the range (a looping sequence shows the same few frames
            # over and over): the layer's batch keys say which frames share one
            keys = item.layer.get_keys(layer_times[idx])
            fetched: dict = {}
            for j, f in enumerate(idx):
                k = keys[j]
                if k not in fetched:
                    fetched[k] = item.fg_image_device(float(times[f]), comp.cache)
                images[j] = fetched[k]
        else:
            for j, f in enumerate(idx):
                images[j] = item.fg_image_device(float(times[f]), comp.cache)
        for j, image in enumerate(images):
            if image is None:
                continue
            by_size.setdefault((image.shape[1], image.shape[0]), []).append(j)
        for size, js in by_size.items():
            js_arr = np.asarray(js)
            sub = type(samples)(samples.anchor_point[js_arr], samples.position[js_arr], samples.scale[js_arr],
                                samples.rotation[js_arr], samples.opacity[js_arr], samples.origin_point,
                                samples.blending_mode)
            plan = plan_affine(sub, size, L)
            # one vectorised fill per size; what differs from frame to frame besides the transform is the
            # image (pointer, row stride, whether its alpha is known to be 255 throughout)
            group = np.zeros(len(js), dtype=_native.LAYER_DTYPE)
            fill_descriptors(group, images[js[0]], plan, sub.opacity, samples.blending_mode)
            group["src"] = [images[j].data_ptr() for j in js]
            group["src_stride"] = [images[j].stride(0) for j in js]
            opaque = np.fromiter((bool(getattr(images[j], "mvb_opaque", False)) for j in js), dtype=bool, count=len(js))
            group["flags"] = (group["flags"] & ~_native.MVB_LAYER_OPAQUE_SOURCE) | np.where(opaque, _native.MVB_LAYER_OPAQUE_SOURCE, 0)
            maps = [getattr(images[j], "mvb_alpha_map", None) for j in js]
            group["alpha_map"] = [0 if m is None else m.data_ptr() for m in maps]
            rows[js_arr] = group
            ok[js_arr] = plan.valid
            keep.extend(images[j] for j in js)
        per_layer.append((idx[ok], rows[ok]))

    if static_jobs:
        # The affine set-up of every static layer in one call: its cost is a fixed number of small numpy
        # calls whatever the number of samples, so 16 layers x 150 frames cost what one layer does.
        from .enum import Direction
        lens = [len(job[1]) for job in static_jobs]
        cat = lambda field: np.concatenate([getattr(job[4], field) for job in static_jobs])   # noqa: E731
        sizes = [(job[3].shape[1], job[3].shape[0]) for job in static_jobs]
        origin = [Direction.to_vector(job[4].origin_point, (float(w), float(h))) for job, (w, h) in zip(static_jobs, sizes)]
        rep = lambda vals: np.repeat(np.asarray(vals, dtype=np.float64), lens)   # noqa: E731
        plan = plan_affine_xy(cat("anchor_point"), cat("position"), cat("scale"), cat("rotation"),
                              rep([o[0] for o in origin]), rep([o[1] for o in origin]),
                              rep([sz[0] for sz in sizes]), rep([sz[1] for sz in sizes]), L)
        start = 0
        for (slot, idx, inside, image, samples), m in zip(static_jobs, lens):
            one = type(plan)(plan.matrix[start:start + m], plan.bbox[start:start + m], plan.valid[start:start + m])
            rows = np.zeros(m, dtype=_native.LAYER_DTYPE)
            fill_descriptors(rows, image, one, samples.opacity, samples.blending_mode)
            sel = inside & one.valid
            per_layer[slot] = (idx[sel], rows[sel])
            start += m

    counts = np.zeros(n, dtype=np.int64)
    for frames, _ in per_layer:
        np.add.at(counts, frames, 1)
    offsets = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    table = np.zeros(int(offsets[-1]), dtype=_native.LAYER_DTYPE)
    cursor = offsets[:-1].astype(np.int64).copy()
    for frames, rows in per_layer:  # paint order == layer order within each frame
        table[cursor[frames]] = rows
        cursor[frames] += 1
    return RangePlan(table, offsets, keep)


def launch_plan(plan: RangePlan, out: torch.Tensor, bg_color) -> None:
    """Enqueue the fused kernel for every frame of ``plan`` writing ``out[f]``."""
    n, H, W, _ = out.shape
    assert n == plan.n_frames
    bg = np.asarray(bg_color, dtype=np.uint8).reshape(4)
    base, step = out.data_ptr(), out.stride(0)
    frame_ptrs = (base + step * np.arange(n, dtype=np.int64)).astype(np.uint64)
    _native.check(_native.lib().mvb_composite_frames(
        n, frame_ptrs.ctypes.data, W, H, out.stride(1), bg.ctypes.data, plan.offsets.ctypes.data,
        plan.descriptors.ctypes.data if len(plan.descriptors) else None, 0, stream_ptr()))


def render_frames(comp: Composition, times: Sequence[float], bg_color=(0, 0, 0, 255), out: torch.Tensor | None = None,
                  plan_hook=None):
    """Render every time in ``times`` into a CUDA tensor ``(N, H // L, W // L, 4)``.

    Times outside ``[0, duration)`` are not allowed here (``Composition.__call__`` returns None
    for them).  Asynchronous: returns after the launches are enqueued on the current stream.
    ``plan_hook(plan, first_frame_index) -> plan``, if given, sees every chunk's ``RangePlan`` before
    it is launched (instrumentation: e.g. the benchmark redirects descriptors to twin copies of the
    sources)."""
    require_cuda()
    times = np.asarray(times, dtype=np.float64)
    assert times.ndim == 1
    assert np.all((times >= 0.0) & (times < comp.duration)), "all times must lie in [0, duration)"
    H, W = comp.frame_shape()
    if out is None:
        out = torch.empty((len(times), H, W, 4), dtype=torch.uint8, device="cuda")
    else:
        assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (len(times), H, W, 4)
        assert out.stride(3) == 1 and out.stride(2) == 4
    if len(times) == 0:
        return out
    # Planned and enqueued in chunks: the device renders chunk i while the host evaluates the
    # keyframes of chunk i + 1 (launches are asynchronous), so a batch costs max(host, device)
    # instead of their sum.
    step = plan_chunk(comp)
    for i in range(0, len(times), step):
        plan = plan_range(comp, times[i:i + step])
        if plan_hook is not None:
            plan = plan_hook(plan, i)
        launch_plan(plan, out[i:i + step], bg_color)
    return out


def stream_frames(comp: Composition, times: Sequence[float], bg_color=(0, 0, 0, 255), batch: int = 8,
                  copy: bool = True) -> Iterator[np.ndarray]:
    """Yield host uint8 frames for ``times`` in order.

    With ``copy=False`` the yielded array is a view into a pinned staging buffer and is valid ONLY
    until the next frame is requested from the generator (asking for the next frame may start the
    copy that overwrites it).  That is enough for a writer that consumes each frame immediately, like
    the reference's ``writer.append_data`` loop; keep ``copy=True`` to hold on to frames.

    Two device batches and two pinned host batches alternate: while batch i is copied to the host on
    a side stream (one cudaMemcpyAsync per batch) and consumed by the caller, batch i + 1 is planned
    and rendered."""
    require_cuda()
    times = np.asarray(times, dtype=np.float64)
    H, W = comp.frame_shape()
    if len(times) == 0:
        return
    batch = max(1, min(batch, len(times)))
    ring = _staging_ring(batch, H, W)
    try:
        yield from _stream_with(ring, comp, times, bg_color, batch, copy)
    finally:
        ring["busy"] = False


_RINGS: dict = {}


def _staging_ring(batch: int, H: int, W: int) -> dict:
    """Two device batches, two pinned host batches, a copy stream and their events, kept per (device,
    batch, frame size) between calls: pinning half a gigabyte takes ~0.1-1 s, far more than rendering
    the frames that go through it.  A ring in use by a live generator is not shared (a second concurrent
    stream of the same shape gets a private one)."""
    key = (torch.cuda.current_device(), batch, H, W)
    ring = _RINGS.get(key)
    if ring is not None and not ring["busy"]:
        for e in ring["copy_done"]:
            e.synchronize()   # an abandoned generator may have left a copy in flight
        ring["busy"] = True
        return ring
    fresh = {
        "dev": [torch.empty((batch, H, W, 4), dtype=torch.uint8, device="cuda") for _ in range(2)],
        "host": [torch.empty((batch, H, W, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)],
        "copy_stream": torch.cuda.Stream(),
        "render_done": [torch.cuda.Event() for _ in range(2)],
        "copy_done": [torch.cuda.Event() for _ in range(2)],
        "busy": True,
    }
    if ring is None:
        if len(_RINGS) >= 4:   # a handful of shapes at most (full size and draft levels)
            _RINGS.pop(next(iter(_RINGS)))
        _RINGS[key] = fresh
    return fresh


def _stream_with(ring: dict, comp: Composition, times: np.ndarray, bg_color, batch: int, copy: bool) -> Iterator[np.ndarray]:
    dev, host, copy_stream = ring["dev"], ring["host"], ring["copy_stream"]
    render_done, copy_done = ring["render_done"], ring["copy_done"]
    main = torch.cuda.current_stream()

    # Keyframes are evaluated for plan_chunk(comp) frames at a time (a handful of numpy calls per layer
    # whatever the count), independently of the copy batch: a batch takes its slice of the current plan.
    # The first plan is short, so the first frames leave early and the long plans that follow are made
    # while copies are in flight.
    step = max(batch, plan_chunk(comp) // batch * batch)
    bounds = [0, min(2 * batch, len(times))]
    while bounds[-1] < len(times):
        bounds.append(min(bounds[-1] + step, len(times)))
    state = {"chunk": -1, "plan": None}

    def plan_for(first: int, n: int) -> RangePlan:
        c = int(np.searchsorted(bounds, first, side="right")) - 1
        if state["chunk"] != c:
            state["plan"] = plan_range(comp, times[bounds[c]:bounds[c + 1]])
            state["chunk"] = c
        return state["plan"].slice(first - bounds[c], first - bounds[c] + n)

    def submit(slot: int, first: int, n: int) -> None:
        main.wait_event(copy_done[slot])  # the previous copy out of dev[slot] must be finished
        launch_plan(plan_for(first, n), dev[slot][:n], bg_color)
        render_done[slot].record(main)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(render_done[slot])
            host[slot][:n].copy_(dev[slot][:n], non_blocking=True)
            copy_done[slot].record(copy_stream)

    assert np.all((times >= 0.0) & (times < comp.duration)), "all times must lie in [0, duration)"
    starts = list(range(0, len(times), batch))
    counts = [min(batch, len(times) - a) for a in starts]
    submit(0, starts[0], counts[0])
    for i, n in enumerate(counts):
        slot = i & 1
        if i + 1 < len(starts):
            submit(slot ^ 1, starts[i + 1], counts[i + 1])
        copy_done[slot].synchronize()
        frames = host[slot].numpy()
        for j in range(n):
            yield frames[j].copy() if copy else frames[j]

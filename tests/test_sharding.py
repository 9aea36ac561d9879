"""Frame-range sharding across ranks: contiguous, disjoint, exhaustive -- checked
in a real world_size-2 torch.distributed (gloo) job on CPU, plus the host half of range planning."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from movis_b200.render import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 150, 1800):
        for g in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, g) for r in range(g)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_weighted_shard_range_partitions():
    """Uneven contiguous ranges (ranks that deliver frames faster own more of the timeline)."""
    for n in (0, 1, 7, 150, 1800):
        for w in ([1.0], [1.0, 1.0], [9.0, 9.0, 9.0, 8.0, 18.0, 18.0, 18.0, 18.0], [0.2, 5.0, 1.0]):
            g = len(w)
            parts = [shard_range(n, r, g, w) for r in range(g)]
            assert [i for p in parts for i in p] == list(range(n))
            for r in range(g):   # within one frame of the exact share
                assert abs(len(parts[r]) - n * w[r] / sum(w)) <= 1.0 + 1e-9
    assert [len(shard_range(150, r, 2, [1.0, 1.0])) for r in range(2)] == [75, 75]
    assert [len(shard_range(1200, r, 8, [9, 9, 9, 9, 18, 18, 18, 18])) for r in range(8)] == [100] * 4 + [200] * 4


def test_cost_balanced_shard_range():
    """Contiguous ranges that balance a per-frame cost (and combine with per-rank weights)."""
    rng = np.random.default_rng(0)
    for n in (1, 7, 150, 1200):
        for g in (1, 2, 4, 8):
            cost = rng.uniform(0.5, 2.0, n) * np.linspace(1.0, 1.5, n)
            parts = [shard_range(n, r, g, costs=cost) for r in range(g)]
            assert [i for p in parts for i in p] == list(range(n))
            if n >= 150:
                sums = [cost[p.start:p.stop].sum() for p in parts]
                assert max(sums) - min(sums) <= 2 * cost.max() + 1e-9
    n, cost = 1200, np.linspace(1.0, 1.4, 1200)
    parts = [shard_range(n, r, 8, costs=cost) for r in range(8)]
    assert len(parts[0]) > len(parts[-1])                      # cheap frames: longer range
    both = [shard_range(n, r, 2, weights=[1.0, 3.0], costs=np.ones(n)) for r in range(2)]
    assert [len(p) for p in both] == [300, 900]
    import movis_b200 as mv
    from movis_b200 import scenes
    from movis_b200.render import frame_costs
    c = frame_costs(scenes.build_s2(mv, div=24), np.arange(0, 5, 1 / 30))
    assert c.shape == (150,) and c[-1] > c[0] > 0              # the layers grow over the timeline


def test_range_plan_slices_and_batch_keys_need_no_gpu():
    """Host-only pieces of the range renderer: RangePlan.slice (what stream_frames launches per copy
    batch) and the batch key lookups (TimelineMixin.get_states / ImageSequence.get_keys, scenes.LoopedLayer)."""
    import movis_b200 as mv
    from movis_b200 import _native, scenes
    from movis_b200.render import RangePlan
    desc = np.zeros(10, dtype=_native.LAYER_DTYPE)
    desc["bbox_x"] = np.arange(10)
    plan = RangePlan(desc, np.array([0, 3, 3, 7, 10], dtype=np.int32), [])
    part = plan.slice(1, 3)
    assert part.n_frames == 2 and part.offsets.tolist() == [0, 0, 4] and part.descriptors["bbox_x"].tolist() == [3, 4, 5, 6]
    assert plan.slice(0, 4).offsets.tolist() == plan.offsets.tolist() and plan.slice(3, 4).descriptors["bbox_x"].tolist() == [7, 8, 9]
    frames = [np.full((2, 2, 4), i, np.uint8) for i in range(4)]
    seq = mv.layer.ImageSequence.from_files(frames, each_duration=0.25)
    ts = np.array([-0.1, 0.0, 0.24, 0.25, 0.6, 0.99, 1.0, 3.0])
    assert seq.get_keys(ts).tolist() == [seq.get_key(float(t)) for t in ts] == [-1, 0, 0, 1, 2, 3, -1, -1]
    loop = scenes.LoopedLayer(seq, 2.0)
    ts = np.array([-0.5, 0.0, 0.3, 1.0, 1.26, 1.99, 2.0])
    assert loop.get_keys(ts).tolist() == [loop.get_key(float(t)) for t in ts] == [None, 0, 1, 0, 1, 3, None]


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, n_frames: int) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import movis_b200 as mv
        from movis_b200 import scenes
        mine = shard_range(n_frames, rank, world)
        # every rank evaluates the keyframes of ITS frames only; the union must equal a
        # single-process evaluation (frames are independent: no data-path collective)
        comp = scenes.build_s2(mv, div=24)
        times = np.arange(0, 5, 1 / 30)[:n_frames]
        item = comp["l3"]
        local = item.transform.sample(times[mine.start:mine.stop] - item.offset)
        full = item.transform.sample(times - item.offset)
        assert np.array_equal(local.rotation, full.rotation[mine.start:mine.stop])
        assert np.array_equal(local.opacity, full.opacity[mine.start:mine.stop])
        owned = torch.zeros(n_frames, dtype=torch.int64)
        owned[mine.start:mine.stop] = 1
        dist.all_reduce(owned)  # test-only bookkeeping, not part of the render path
        assert bool((owned == 1).all())
        t = torch.tensor([float(len(mine))], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks reduction bench.py uses
        assert t.item() == (n_frames + world - 1) // world
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_job():
    mp.spawn(_worker, args=(2, _free_port(), 37), nprocs=2, join=True)

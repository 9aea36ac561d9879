"""BASELINE.json configs[4]: the 60 s 4K 30 fps timeline (1800 frames, 24 layers + effects), frames
sharded by time index across the GPUs of one box.

    python tools/bench_s5.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29533 tools/bench_s5.py                  # N GPUs

Rank r of G renders the contiguous frame range [floor(r N / G), floor((r + 1) N / G))
(render.shard_range) and delivers uint8 frames to pinned host memory (render.stream_frames, the loop
of Composition._write_video).  No pixels cross GPUs; torch.distributed is used for the barrier and
the max-over-ranks time only.  Prints one JSON line on rank 0: delivered frames/s (whole job) and
device-only frames/s (frames left in HBM)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import movis_b200 as mv  # noqa: E402
from movis_b200 import scenes  # noqa: E402
from movis_b200.contrib.segmentation import ChromaKey  # noqa: E402
from movis_b200.render import render_frames, shard_range, stream_frames  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


N, FPS = 1800, 30.0
comp = scenes.build_s5(mv, ChromaKey)
mine = shard_range(N, rank, world)
times = np.asarray(mine, dtype=np.float64) / FPS
bg = (0, 0, 0, 255)
for _ in stream_frames(comp, times[:16], bg_color=bg, batch=8, copy=False):   # warm-up: uploads, effect caches
    pass
barrier()
t0 = time.perf_counter()
n = 0
for _ in stream_frames(comp, times, bg_color=bg, batch=10, copy=False):
    n += 1
barrier()
delivered_s = max_over_ranks(time.perf_counter() - t0)

# device only: the same range in batches of 60 frames, frames stay in HBM
out = torch.empty((60,) + comp.frame_shape() + (4,), dtype=torch.uint8, device="cuda")
barrier()
t0 = time.perf_counter()
for i in range(0, len(times), 60):
    chunk = times[i:i + 60]
    render_frames(comp, chunk, bg_color=bg, out=out[:len(chunk)])
barrier()
device_s = max_over_ranks(time.perf_counter() - t0)
if rank == 0:
    print(json.dumps({"scene": "S5: 60 s 4K 30 fps, 24 layers + effects", "n_gpus": world, "frames": N,
                      "delivered_frames_per_s": N / delivered_s, "device_frames_per_s": N / device_s,
                      "frames_per_rank": len(mine)}))
if world > 1:
    dist.destroy_process_group()

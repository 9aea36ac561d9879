"""Ceiling of delivered frames on a multi-GPU box: a plain pinned device -> host copy (and host -> device)
running on N ranks AT ONCE, one rank per GPU, optionally with each rank bound to its GPU's NUMA node
(movis_b200.device.bind_host_to_gpu).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29544 tools/pcie_concurrent.py [--bind] [--mib 512] [--seconds 2]

Rank 0 prints one JSON line: per-rank and aggregate GB/s for both directions (max-over-ranks time)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from movis_b200.device import bind_host_to_gpu, gpu_numa_node  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bind", action="store_true")
ap.add_argument("--mib", type=int, default=512)
ap.add_argument("--seconds", type=float, default=2.0)
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
binding = bind_host_to_gpu(local) if args.bind else {"numa_node": gpu_numa_node(local), "cpus": [], "mempolicy": False}
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


n = args.mib << 20
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h.fill_(1)
result = {"n_ranks": world, "bind": args.bind, "mib_per_copy": args.mib,
          "host_cpus": len(os.sched_getaffinity(0)), "numa_node_rank0": binding["numa_node"]}
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    reps = max(2, int(args.seconds / max(time.perf_counter() - t0, 1e-4)))
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    mine = reps * n / (time.perf_counter() - t0) / 1e9
    rates = torch.tensor([mine], dtype=torch.float64, device="cuda")
    if world > 1:
        gathered = [torch.zeros_like(rates) for _ in range(world)]
        dist.all_gather(gathered, rates)
        per_rank = [float(g.item()) for g in gathered]
    else:
        per_rank = [mine]
    result[name + "_GBps_per_rank"] = [round(v, 2) for v in per_rank]
    result[name + "_GBps_aggregate"] = round(sum(per_rank), 2)
    barrier()
if rank == 0:
    print(json.dumps(result))
if world > 1:
    dist.destroy_process_group()

"""Host->device bandwidth ceiling of this box, measured the way bench.py's e2e path uses it: every rank copies a
pinned buffer to its own GPU, all ranks at once (nvbandwidth-style host_to_device_memcpy, run under torchrun).
    python -m torch.distributed.run --nproc-per-node N tools/h2d_ceiling.py [MiB per copy]
Rank 0 prints one JSON line: per-rank GB/s when alone (rank 0 only) and the aggregate with all ranks copying."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = mib * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")


def run(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return n * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(3)
alone = None
if world > 1:
    dist.barrier()
    if rank == 0:
        alone = run(20)
    dist.barrier()
else:
    alone = run(20)
together = run(40)
if world > 1:
    t = torch.tensor([together], device="cuda")
    dist.all_reduce(t)
    tot = float(t)
    mn = torch.tensor([together], device="cuda")
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
else:
    tot, mn = together, torch.tensor([together])
if rank == 0:
    print(json.dumps({"n_gpus": world, "MiB_per_copy": mib, "GBps_rank0_alone": alone, "GBps_aggregate_all_ranks": tot,
                      "GBps_slowest_rank": float(mn), "host_cpus": os.cpu_count()}), flush=True)
if world > 1:
    dist.destroy_process_group()

"""Batch sharding across the GPUs of one box.

Hands are independent, so the data path needs no collective: each rank owns a contiguous slice of
the batch and its own replica of the 1.5 MB MANO constants.  The only exchange per step is one
all-reduce(sum) of a packed fp32 record [loss_sum, abs_sum, mask_count, n_hands] so every rank
can report the global loss; per-hand parameter gradients stay on the rank that owns the hand.
The helpers work on any torch.distributed backend (NCCL on the box, gloo in the CPU tests).
"""
from __future__ import annotations

import datetime
import os

import torch
import torch.distributed as dist

PACK = 4   # loss_sum, abs_sum, mask_count, n_hands


def shard_bounds(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of ``total`` hands for ``rank``; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group
    when WORLD_SIZE > 1 (127.0.0.1 rendezvous is whatever MASTER_ADDR says)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                timeout=datetime.timedelta(seconds=180))
    return rank, world, local


def pack_totals(loss_weighted_sum: torch.Tensor, abs_sum: torch.Tensor, count: torch.Tensor,
                n_hands: int) -> torch.Tensor:
    out = torch.empty(PACK, dtype=torch.float32, device=loss_weighted_sum.device)
    out[0] = loss_weighted_sum
    out[1] = abs_sum
    out[2] = count
    out[3] = float(n_hands)
    return out


def allreduce_totals(packed: torch.Tensor, group=None, async_op: bool = False):
    """Sum the packed record over ranks in place; a no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        return dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return None


def global_loss(packed: torch.Tensor) -> torch.Tensor:
    """Batch-mean m2d loss over all ranks: each rank contributes local_mean * n_local."""
    return packed[0] / packed[3]


class TotalsReducer:
    """Per-step all-reduce of a FitStep's ``totals`` record that never stalls the compute stream:
    the record is copied into one of two rotating buffers and reduced with ``async_op=True`` (NCCL's
    own stream), so the collective of step i overlaps the kernels of step i+1.  ``totals[3]`` is the
    un-normalised loss (loss * B_local), hence global loss = reduced[3] / total_hands."""

    def __init__(self, device, total_hands: int):
        self.bufs = [torch.zeros(PACK, dtype=torch.float32, device=device) for _ in range(2)]
        self.handles = [None, None]
        self.total_hands = total_hands
        self.i = 0

    def submit(self, totals: torch.Tensor):
        k = self.i & 1
        if self.handles[k] is not None:
            self.handles[k].wait()                 # buffer k was last used two steps ago
        self.bufs[k].copy_(totals, non_blocking=True)
        self.handles[k] = allreduce_totals(self.bufs[k], async_op=True)
        self.i += 1

    def finish(self) -> float:
        """Wait for everything in flight; returns the global loss of the last submitted step."""
        for h in self.handles:
            if h is not None:
                h.wait()
        last = self.bufs[(self.i - 1) & 1]
        return float(last[3].item()) / self.total_hands


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin the calling process to the CPU cores next to its GPU (NVML's ideal affinity) so that pinned
    host buffers are first-touched on the GPU's own NUMA node.  With one process per GPU all pulling
    host memory at PCIe rate, remote-node buffers are what saturates first.  Best effort: returns
    False when NVML or the affinity call is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = local_rank
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                index = int(ids[local_rank])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False

"""Host-side multi-GPU logic on CPU: batch sharding and the packed loss all-reduce over a
world_size-2 gloo group (the NCCL path on the box runs the same code)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dsf_b200 import dist as D


def test_shard_bounds_cover_the_batch_exactly():
    for total in (1, 7, 128, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env(backend="gloo")
    lo, hi = D.shard_bounds(total, r, w)
    per_hand = torch.arange(total, dtype=torch.float32)[lo:hi] * 0.01      # stand-in per-hand losses
    n = hi - lo
    packed = D.pack_totals(per_hand.mean() * n, per_hand.sum(), torch.tensor(float(n * 3)), n)
    D.allreduce_totals(packed)
    t = D.max_over_ranks(float(rank + 1), torch.device("cpu"))
    # the overlapped per-step reducer: totals = [loss, sum, count, loss * n_local]
    red = D.TotalsReducer(torch.device("cpu"), total)
    for it in range(3):
        red.submit(torch.tensor([per_hand.mean(), per_hand.sum(), float(n * 3), per_hand.mean() * n]) * (it + 1))
    out[rank] = (packed.tolist(), D.global_loss(packed).item(), t, red.finish())
    dist.destroy_process_group()


def test_packed_allreduce_world2_gloo():
    world, total = 2, 37
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), total, out), nprocs=world, join=True)
    ref = torch.arange(total, dtype=torch.float32) * 0.01
    for r in range(world):
        packed, loss, t, red_loss = out[r]
        assert abs(red_loss - 3 * ref.mean().item()) < 1e-5
        assert abs(packed[1] - ref.sum().item()) < 1e-4
        assert packed[2] == total * 3 and packed[3] == total
        assert abs(loss - ref.mean().item()) < 1e-5
        assert t == 2.0

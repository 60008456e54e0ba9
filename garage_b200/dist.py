"""Multi-GPU plumbing of the block path: blocks are independent
(src/api/s3/put.rs:134,448), so rank r of R owns a contiguous block range and nothing crosses
GPUs on the data path.  The only collectives are one broadcast of the control struct
{k, m, parity matrix, per-rank ranges} from rank 0 and a MAX/SUM reduction of timings/counters.
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def partition_blocks(n_total: int, world: int):
    """contiguous, disjoint, exhaustive ranges [lo, hi) per rank; sizes differ by at most 1"""
    base, rem = divmod(n_total, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def broadcast_control(k, m, matrix, ranges, device, dist):
    """rank 0 passes (matrix m x k uint8, ranges list); every rank returns (matrix, ranges).
    One int64 tensor [k, m, world, lo0, hi0, ..., P...] so a single broadcast carries it all."""
    import torch

    world = dist.get_world_size()
    n = 3 + 2 * world + k * m
    buf = torch.zeros(n, dtype=torch.int64, device=device)
    if dist.get_rank() == 0:
        flat = [k, m, world]
        for lo, hi in ranges:
            flat += [lo, hi]
        flat += [int(v) for v in np.asarray(matrix, dtype=np.uint8).reshape(-1)]
        buf.copy_(torch.tensor(flat, dtype=torch.int64))
    dist.broadcast(buf, 0)
    h = buf.cpu().numpy()
    if int(h[0]) != k or int(h[1]) != m or int(h[2]) != world:
        raise RuntimeError("control broadcast mismatch: rank 0 has (k,m,world)=%s" % (h[:3],))
    rg = [(int(h[3 + 2 * r]), int(h[4 + 2 * r])) for r in range(world)]
    P = h[3 + 2 * world:].astype(np.uint8).reshape(m, k)
    return P, rg


def reduce_max(value: float, device, dist) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: int, device, dist) -> int:
    import torch

    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())

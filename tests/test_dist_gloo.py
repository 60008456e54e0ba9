"""CPU-only, world_size 2 over gloo: the N>1 host logic (range partition, control broadcast of
the generator matrix, max-over-ranks / sum reductions) that bench.py runs over NCCL on GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from garage_b200 import dist as D  # noqa: E402


def test_partition_blocks():
    for n, w in ((65536, 8), (4096, 1), (10, 3), (7, 8), (0, 2)):
        r = D.partition_blocks(n, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 1
    assert D.partition_blocks(65536, 8)[3] == (24576, 32768)  # BASELINE config 4: 8192 per GPU


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import oracle_lib as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, m, n_total = 10, 4, 65536 + 3
        dev = torch.device("cpu")
        if rank == 0:
            P0 = O.build_matrix(k, m, 0)  # stands in for garage_ec_matrix() on the GPU box
            ranges0 = D.partition_blocks(n_total, world)
        else:
            P0, ranges0 = None, None
        P, ranges = D.broadcast_control(k, m, P0, ranges0, dev, dist)
        mx = D.reduce_max(10.0 + rank, dev, dist)
        sm = D.reduce_sum(rank + 1, dev, dist)
        # every rank regenerates ITS blocks from the shared counter stream: disjoint, deterministic
        lo, hi = ranges[rank]
        first = O.fill_random(64, 0x6761726167650010, lo * (1 << 20))
        q.put((rank, P.tolist(), ranges, mx, sm, first.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_control_broadcast_world2():
    import torch.multiprocessing as mp

    import oracle_lib as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs])
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    want = O.build_matrix(10, 4, 0).tolist()
    (r0, P0, rg0, mx0, sm0, f0), (r1, P1, rg1, mx1, sm1, f1) = res
    assert (r0, r1) == (0, 1)
    assert P0 == want and P1 == want
    assert rg0 == rg1 == D.partition_blocks(65536 + 3, 2)
    assert mx0 == mx1 == 11.0 and sm0 == sm1 == 3
    assert f0 != f1  # different block ranges -> different bytes

#!/usr/bin/env python
"""What limits the HOST-buffer (e2e) step on a multi-GPU host?  Run under torchrun with N ranks: the same e2e
step as bench.py (garage_ec_encode + garage_ec_reconstruct, HOST mode, NUMA-local pinned buffers) is timed with
different SUBSETS of ranks active (the others wait at the barrier) and with plain vs write-combined pinned
memory.  If a socket's four GPUs slow each other down while GPUs of different sockets do not, the per-socket
memory / IO system is the limiter; if write-combined buffers lift the rate, coherence traffic is part of it.

    python -m torch.distributed.run --nproc-per-node 8 tools/e2e_probe.py [--blocks 2048]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import garage_b200 as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=2048)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
k, m, B, n = 10, 4, 1 << 20, args.blocks
tot = k + m
enc, dec = G.GarageEc(local, k, m), G.GarageEc(local, k, m)
enc.bind_thread()
L = enc.shard_len(B)
stride = enc.stride_for(L)
data = torch.empty(n * k * stride, dtype=torch.uint8, device=dev)
enc.fill_random(data, n * k * stride, 5 + rank, 0)
data.view(n, k, stride)[:, :, L:] = 0
lens_d = torch.full((n,), L, dtype=torch.int32, device=dev)
par = torch.zeros(n * m * stride, dtype=torch.uint8, device=dev)
enc.encode(data, par, stride, n, shard_len=lens_d)
shards = torch.cat([data.view(n, k, stride), par.view(n, m, stride)], dim=1).contiguous()
present = np.ones((n, tot), dtype=np.uint8)
rng = np.random.default_rng(rank)
for s in range(n):
    present[s, rng.choice(tot, m, replace=False)] = 0
h_lens = np.full(n, L, dtype=np.uint32)
h_status = np.zeros(n, dtype=np.int32)
results = {}


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for wc in (False, True):
    h_data, p1 = enc.host_alloc(n * k * stride, write_combined=wc)
    h_par, p2 = enc.host_alloc(n * m * stride, write_combined=wc)
    h_sh, p3 = enc.host_alloc(n * tot * stride, write_combined=wc)
    torch.from_numpy(h_data).copy_(data)
    torch.from_numpy(h_sh).copy_(shards.view(-1))
    torch.cuda.synchronize()
    subsets = [list(range(world))]
    if world >= 8:
        subsets += [[0, 1, 2, 3], [0, 1, 4, 5], [0, 4], [0, 1], [0]]
    elif world >= 2:
        subsets += [[0]]
    for sub in subsets:
        active = rank in sub
        if active:
            enc.encode(h_data, h_par, stride, n, shard_len=h_lens)  # warm
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(args.steps):
                enc.encode(h_data, h_par, stride, n, shard_len=h_lens)
                dec.reconstruct(h_sh, present, stride, n, status=h_status, shard_len=h_lens)
        el = time.perf_counter() - t0
        t = torch.tensor([el if active else 0.0], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        if world > 1:
            dist.all_gather(allt, t)
        else:
            allt = [t]
        barrier()
        per = [2 * n * B * args.steps / float(x.item()) / 2**30 if float(x.item()) > 0 else 0.0 for x in allt]
        results["%s ranks=%s" % ("write-combined" if wc else "plain pinned", ",".join(map(str, sub)))] = {
            "per_rank_GiBs": [round(x, 1) for x in per if x > 0],
            "sum_GiBs": round(2 * n * B * args.steps * len(sub) / max(float(x.item()) for x in allt) / 2**30, 1)}
    for p in (p1, p2, p3):
        enc.host_free(p)
if rank == 0:
    print(json.dumps({"blocks_per_rank": n, "gpu_numa_node_rank0": enc.numa_info()[0], "results": results}, indent=1))
if world > 1:
    dist.destroy_process_group()

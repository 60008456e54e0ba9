#!/usr/bin/env python
"""BASELINE config 5: scrub/repair sweep, mixed 50/50 RS(6,3) / RS(10,4) stripes of 1 MiB blocks,
every shard independently corrupted with p = 0.10; throughput = stripe payload bytes
healed-or-verified per second (device-resident, per GPU).  Run under torchrun for N GPUs.

    python tools/sweep_bench.py [--stripes 2048] [--iters 5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import garage_b200 as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--stripes", type=int, default=2048, help="stripes per code per GPU")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--p", type=float, default=0.10)
ap.add_argument("--sum-kind", type=int, default=1, help="per-shard tag: 1 adler8 (default), 0 blake2sum")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
B = 1 << 20
n = args.stripes
res = {}
total_payload, total_ms = 0, 0.0
for (k, m) in ((6, 3), (10, 4)):
    tot = k + m
    ec = G.GarageEc(local, k, m)
    ec.set_sum_kind(args.sum_kind)
    L = ec.shard_len(B)
    stride = ec.stride_for(L)
    data = torch.empty(n * k * stride, dtype=torch.uint8, device="cuda")
    ec.fill_random(data, n * k * stride, 0x6761726167650010 + rank, 0)
    data.view(n, k, stride)[:, :, L:] = 0
    lens = torch.full((n,), L, dtype=torch.int32, device="cuda")
    par = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
    ec.encode(data, par, stride, n, shard_len=lens)
    shards = torch.cat([data.view(n, k, stride), par.view(n, m, stride)], dim=1).contiguous()
    del data, par
    sums = torch.zeros(n * tot * 32, dtype=torch.uint8, device="cuda")
    ec.shard_sums(shards.view(-1), sums, stride, n, tot, shard_len=lens)
    orig = shards.clone()
    g = torch.Generator().manual_seed(99 + rank)
    hit = (torch.rand(n, tot, generator=g) < args.p).cuda()
    pos = torch.randint(0, L, (n, tot), generator=g).cuda()
    bad = torch.zeros(n * tot, dtype=torch.uint8, device="cuda")
    status = torch.zeros(n, dtype=torch.int32, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    per_iter = []
    for it in range(args.iters + 1):
        shards.copy_(orig)
        sidx, iidx = torch.nonzero(hit, as_tuple=True)
        shards[sidx, iidx, pos[sidx, iidx]] ^= 0x5A  # flip one byte in every hit shard
        torch.cuda.synchronize()
        ev0.record()
        ec.scrub_repair(shards.view(-1), sums, bad, stride, n, status=status, shard_len=lens)
        ev1.record()
        torch.cuda.synchronize()
        if it:
            ms += ev0.elapsed_time(ev1)
            per_iter.append(round(ev0.elapsed_time(ev1), 3))
    ms /= args.iters
    nbad = hit.sum(dim=1)
    unrec = int((nbad > m).sum())
    assert torch.equal(bad.view(n, tot).bool(), hit)
    assert int((status != 0).sum()) == unrec
    ok = status == 0
    assert torch.equal(shards[ok], orig[ok])
    # the tag pass alone (detect), on the healed shards
    tag_ms = 0.0
    bad2 = torch.zeros_like(bad)
    for it in range(args.iters + 1):
        torch.cuda.synchronize()
        ev0.record()
        ec.check_sums(shards.view(-1), sums, bad2, stride, n, tot, shard_len=lens)
        ev1.record()
        torch.cuda.synchronize()
        if it:
            tag_ms += ev0.elapsed_time(ev1)
    tag_ms /= args.iters
    res["rs%d_%d" % (k, m)] = {"stripes": n, "ms": round(ms, 3), "per_iter_ms": per_iter, "tag_pass_ms": round(tag_ms, 3),
                               "tag_pass_GBs": round(n * tot * L / tag_ms / 1e6, 1), "payload_GiBs": round(n * B / ms / 1e-3 / 2**30, 1),
                               "corrupt_shards": int(hit.sum()), "stripes_healed": int(((nbad > 0) & (nbad <= m)).sum()),
                               "unrecoverable": unrec}
    total_payload += n * B
    total_ms += ms
    ec.close()
t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "config 5: mixed RS(6,3)/RS(10,4), p=%.2f corrupted shards, detect+reconstruct+rewrite" % args.p,
                      "n_gpus": world, "sum_kind": "adler8" if args.sum_kind else "blake2sum", "sweep_GiBs": round(total_payload * world / (float(t.item()) * 1e-3) / 2**30, 1),
                      "detail": res}))
if world > 1:
    dist.destroy_process_group()

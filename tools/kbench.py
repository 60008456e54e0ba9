#!/usr/bin/env python
"""Kernel micro-bench for tuning: times garage_ec_encode / _reconstruct / _verify (DEVICE mode)
of one .so variant on the BASELINE config-2/3 workload.  Not the contract bench (bench.py).

    python tools/kbench.py [--so path/to/libgarage_ec.so] [--k 10 --m 4 --blocks 4096 --iters 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--so", default=None)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--m", type=int, default=4)
ap.add_argument("--blocks", type=int, default=4096)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--erasures", type=int, default=None)
ap.add_argument("--same-pattern", action="store_true", help="all stripes lose the same shards")
ap.add_argument("--tag", default="")
ap.add_argument("--stride-pad", type=int, default=0, help="extra bytes between shards (multiple of 16)")
args = ap.parse_args()

from garage_b200 import _build  # noqa: E402

if args.so:
    _build.SO = os.path.abspath(args.so)
    _build.stale = lambda: False
import torch  # noqa: E402

import garage_b200 as G  # noqa: E402

k, m, n = args.k, args.m, args.blocks
tot = k + m
e = args.erasures if args.erasures is not None else m
B = 1 << 20
ec = G.GarageEc(0, k, m)
L = ec.shard_len(B)
stride = ec.stride_for(L) + args.stride_pad
shards = torch.zeros(n * tot * stride, dtype=torch.uint8, device="cuda")
sh3 = shards.view(n, tot, stride)
data = torch.empty(n * k * stride, dtype=torch.uint8, device="cuda")
ec.fill_random(data, n * k * stride, 1, 0)
data.view(n, k, stride)[:, :, L:] = 0
lens = torch.full((n,), L, dtype=torch.int32, device="cuda")
parity = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
ec.encode(data, parity, stride, n, shard_len=lens)
sh3[:, :k] = data.view(n, k, stride)
sh3[:, k:] = parity.view(n, m, stride)
orig = shards.clone()
g = torch.Generator().manual_seed(7)
if args.same_pattern:
    erased = torch.arange(e).repeat(n, 1)
else:
    erased = torch.rand(n, tot, generator=g).argsort(dim=1)[:, :e]
present = torch.ones(n, tot, dtype=torch.uint8)
if e:
    present.scatter_(1, erased, 0)
pd = present.cuda()
sh3[~pd.bool()] = 0
status = torch.zeros(n, dtype=torch.int32, device="cuda")
mm = torch.zeros(n, dtype=torch.int32, device="cuda")


def timeit(fn, ctx):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ctx.set_timing(True)
    ctx.timing_read()
    for _ in range(args.iters):
        fn()
    torch.cuda.synchronize()
    ms, cnt = ctx.timing_read()
    ctx.set_timing(False)
    return ms / max(cnt, 1)


t_enc = timeit(lambda: ec.encode(data, parity, stride, n, shard_len=lens), ec)
t_dec = timeit(lambda: ec.reconstruct(shards, pd, stride, n, status=status, shard_len=lens), ec)
ok = torch.equal(shards, orig) and int(status.abs().sum()) == 0
t_ver = timeit(lambda: ec.verify(shards, mm, stride, n, shard_len=lens), ec)
ok = ok and int(mm.abs().sum()) == 0
peak = 6569.3
enc_b = n * (k + m) * L
dec_b = n * (k + e) * L
print(json.dumps({
    "tag": args.tag or (args.so or "default"), "k": k, "m": m, "blocks": n, "erasures": e, "ok": bool(ok),
    "encode_ms": round(t_enc, 4), "encode_GBs": round(enc_b / t_enc / 1e6, 1), "encode_frac": round(enc_b / t_enc / 1e6 / peak, 4),
    "decode_ms": round(t_dec, 4), "decode_GBs": round(dec_b / t_dec / 1e6, 1), "decode_frac": round(dec_b / t_dec / 1e6 / peak, 4),
    "verify_ms": round(t_ver, 4), "verify_GBs": round(enc_b / t_ver / 1e6, 1), "verify_frac": round(enc_b / t_ver / 1e6 / peak, 4),
}))

#!/usr/bin/env python
"""Wall-clock latency of the block-level HOST entry points per batch size (what one dispatcher call of the
batching front-end costs): garage_ec_encode_blocks_with_sums and garage_ec_reconstruct_stripes on pinned
buffers, n = 1 .. 64 blocks of 1 MiB, RS(10,4), adler8 tags."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import garage_b200 as G  # noqa: E402

k, m, B = 10, 4, 1 << 20
tot = k + m
ec = G.GarageEc(0, k, m)
ec.set_sum_kind(G.SUM_ADLER8)
L = ec.shard_len(B)
stride = ec.stride_for(L)
NMAX = 64
blk, pb = ec.host_alloc(NMAX * B)
blk[:] = np.random.default_rng(0).integers(0, 256, NMAX * B, dtype=np.uint8)
par, pp = ec.host_alloc(NMAX * m * stride)
sums, ps = ec.host_alloc(NMAX * tot * 32)
stripes, pst = ec.host_alloc(NMAX * tot * stride)
res = []
for n in (1, 2, 4, 8, 16, 32, 64):
    blocks = [blk[i * B:(i + 1) * B] for i in range(n)]
    for _ in range(3):
        ec.encode_blocks(blocks, par[: n * m * stride], stride, sums_out=sums[: n * tot * 32])
    t0 = time.perf_counter()
    it = 20
    for _ in range(it):
        ec.encode_blocks(blocks, par[: n * m * stride], stride, sums_out=sums[: n * tot * 32])
    t_enc = (time.perf_counter() - t0) / it
    # stripes for reconstruct: data shards from the blocks, parity from the call above, 2 data shards lost
    sv = stripes[: n * tot * stride].reshape(n, tot, stride)
    for s in range(n):
        flat = np.zeros(k * L, dtype=np.uint8)
        flat[:B] = blocks[s]
        sv[s, :k, :L] = flat.reshape(k, L)
        sv[s, k:] = par[: n * m * stride].reshape(n, m, stride)[s]
    keep = sv.copy()
    present = np.ones((n, tot), dtype=np.uint8)
    present[:, [1, 7]] = 0
    want = np.zeros((n, tot), dtype=np.uint8)
    want[:, [1, 7]] = 1
    lens = np.full(n, L, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    views = [sv[s].reshape(-1) for s in range(n)]
    sv[:, [1, 7]] = 0
    for _ in range(3):
        ec.reconstruct_stripes(views, present, stride, want=want, status=st, shard_len=lens)
    assert np.array_equal(sv[:, :, :L], keep[:, :, :L])
    t0 = time.perf_counter()
    for _ in range(it):
        ec.reconstruct_stripes(views, present, stride, want=want, status=st, shard_len=lens)
    t_rec = (time.perf_counter() - t0) / it
    res.append({"n": n, "encode_blocks_ms": round(t_enc * 1e3, 3), "encode_GiBs": round(n * B / t_enc / 2**30, 2),
                "reconstruct_stripes_ms": round(t_rec * 1e3, 3), "reconstruct_GiBs": round(n * B / t_rec / 2**30, 2)})
print(json.dumps(res))

#!/usr/bin/env python
"""Throughput of the C++ BlockManager mirror (row f1: the batching front-end decides real-world
throughput): T client threads PUT 1 MiB blocks (hash + encode + per-shard
sums + store on k+m in-process nodes), then GET them back, healthy and with m nodes down.

    python tools/bm_bench.py [--threads 16] [--blocks 64] [--k 10 --m 4]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from garage_b200 import block_manager as BM  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, default=16)
ap.add_argument("--blocks", type=int, default=64, help="blocks per thread")
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--m", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--linger-us", type=int, default=300)
args = ap.parse_args()

B = 1 << 20
rng = np.random.default_rng(0)
blocks = [[rng.integers(0, 256, B, dtype=np.uint8) for _ in range(args.blocks)] for _ in range(args.threads)]
hashes = [[BM.blake2sum(b) for b in bl] for bl in blocks]  # put.rs:448 (client side, not timed here)
bm = BM.BlockManager(args.k, args.m, batch_max_blocks=args.batch, batch_linger_us=args.linger_us,
                     block_ram_buffer_max=1 << 30)


def run(fn):
    errs = []

    def w(t):
        for i in range(args.blocks):
            if not fn(t, i):
                errs.append((t, i))

    th = [threading.Thread(target=w, args=(t,)) for t in range(args.threads)]
    t0 = time.perf_counter()
    [x.start() for x in th]
    [x.join() for x in th]
    el = time.perf_counter() - t0
    assert not errs, errs[:3]
    return args.threads * args.blocks * B / el / 2**30


def put(t, i):
    return bm.rpc_put_block(hashes[t][i], blocks[t][i]) == BM.OK


def get(t, i):
    rc, got = bm.rpc_get_block(hashes[t][i])
    return rc == BM.OK and got[0] == blocks[t][i][0] and got[-1] == blocks[t][i][-1]


res = {"threads": args.threads, "blocks": args.threads * args.blocks, "k": args.k, "m": args.m}
res["put_GiBs"] = round(run(put), 2)
m0 = bm.metrics()
res["put_batches"], res["avg_put_batch"] = m0["put_batches"], round(m0["put_calls"] / max(m0["put_batches"], 1), 1)
res["encode_call_ms_per_batch"] = round(m0["encode_call_us"] / 1e3 / max(m0["put_batches"], 1), 2)
res["get_healthy_GiBs"] = round(run(get), 2)
for d in range(args.m):
    bm.set_node_up(d, False)
res["get_m_nodes_down_GiBs"] = round(run(get), 2)
m1 = bm.metrics()
res["reconstruct_calls"], res["reconstruct_batches"] = m1["reconstruct_calls"], m1["reconstruct_batches"]
res["reconstruct_call_ms_per_batch"] = round(m1["reconstruct_call_us"] / 1e3 / max(m1["reconstruct_batches"], 1), 2)
print(json.dumps(res))
bm.close()

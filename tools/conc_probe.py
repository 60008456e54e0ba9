#!/usr/bin/env python
"""What does one dispatcher call cost when several dispatchers share the GPU?  T threads, each with its own
garage_ec context and pinned buffers, call garage_ec_encode_blocks_with_sums back to back on batches of n
1 MiB blocks (RS(10,4), adler8 tags).  --scattered: the blocks of a batch are not adjacent in memory (as the
put slots of the block manager are), so the upload is n copies instead of one.

    python tools/conc_probe.py [--threads 1 2 3] [--batch 8 16] [--iters 200]
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import garage_b200 as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--threads", type=int, nargs="+", default=[1, 2, 3])
ap.add_argument("--batch", type=int, nargs="+", default=[8, 16])
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--busy", type=int, default=0, help="extra threads that memcpy 1 MiB buffers in a loop (client load)")
args = ap.parse_args()

k, m, B = 10, 4, 1 << 20
tot = k + m
NMAX = max(args.batch)
TMAX = max(args.threads)


class Worker:
    def __init__(self):
        self.ec = G.GarageEc(0, k, m)
        self.ec.set_sum_kind(G.SUM_ADLER8)
        self.stride = self.ec.stride_for(self.ec.shard_len(B))
        self.blk, self.pb = self.ec.host_alloc(2 * NMAX * B)
        self.blk[:] = 7
        self.out, self.po = self.ec.host_alloc(NMAX * (m * self.stride + tot * 32))

    def args_for(self, n, scattered):
        base = self.blk.ctypes.data
        order = [(2 * i + 1) % (2 * n) if scattered else i for i in range(n)]
        if scattered:
            order = order[::-1]
        ptrs = (C.c_void_p * n)(*[base + o * B for o in order])
        lens = (C.c_uint32 * n)(*([B] * n))
        par = self.out.ctypes.data
        return (self.ec._h, C.cast(ptrs, C.c_void_p), C.cast(lens, C.c_void_p), C.c_size_t(n), C.c_void_p(par),
                C.c_void_p(par + n * m * self.stride), C.c_size_t(self.stride)), (ptrs, lens)


workers = [Worker() for _ in range(TMAX)]
fn = workers[0].ec._L.garage_ec_encode_blocks_with_sums
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
stop = threading.Event()


def busy():
    a = np.zeros(B, dtype=np.uint8)
    b = np.ones(B, dtype=np.uint8)
    while not stop.is_set():
        np.copyto(a, b)


for _ in range(args.busy):
    threading.Thread(target=busy, daemon=True).start()

res = []
for scattered in (False, True):
    for n in args.batch:
        for T in args.threads:
            per = [0.0] * T
            bar = threading.Barrier(T)

            def run(t):
                a, keep = workers[t].args_for(n, scattered)
                for _ in range(5):
                    assert fn(*a) == 0
                bar.wait()
                t0 = time.perf_counter()
                for _ in range(args.iters):
                    fn(*a)
                per[t] = (time.perf_counter() - t0) / args.iters

            th = [threading.Thread(target=run, args=(t,)) for t in range(T)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            ms = sum(per) / T * 1e3
            res.append({"scattered": scattered, "batch": n, "threads": T, "ms_per_call": round(ms, 3),
                        "GiBs_total": round(T * n * B / (ms / 1e3) / 2**30, 1)})
            print(res[-1], flush=True)
stop.set()
print(json.dumps({"busy_threads": args.busy, "batchcopy": os.environ.get("GARAGE_EC_BATCHCOPY", "1"), "results": res}))
for w in workers:
    w.ec.host_free(w.pb)
    w.ec.host_free(w.po)
    w.ec.close()

#!/usr/bin/env python
"""Seconds-long check of the block manager mirror on a GPU box: PUT a few blocks, GET them back healthy and with m
nodes down, rebuild a wiped shard.  No torch import (numpy + ctypes only)."""
import hashlib
import os
import sys
import time

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from garage_b200 import block_manager as BM  # noqa: E402

bm = BM.BlockManager(10, 4)
rng = np.random.default_rng(1)
blocks = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in (1 << 20, 1, 777777, (1 << 20) - 7, 104858)]
hs = [hashlib.blake2b(b).digest()[:32] for b in blocks]  # blake2sum: BLAKE2b-512 truncated (util/data.rs:130-138)
for h, b in zip(hs, blocks):
    assert bm.rpc_put_block(h, np.frombuffer(b, dtype=np.uint8)) == 0
def get(h):
    rc, out = bm.rpc_get_block(h)
    assert rc == 0, rc
    return out.tobytes()


for h, b in zip(hs, blocks):
    assert get(h) == b
    assert bm.storage_nodes_of(h) == [(h[0] % 14 + i) % 14 for i in range(14)]
for d in range(4):
    bm.set_node_up(d, False)
for h, b in zip(hs, blocks):
    assert get(h) == b
for d in range(4):
    bm.set_node_up(d, True)
node = bm.storage_nodes_of(hs[0])[2]
bm.drop_shard(node, hs[0])
assert bm.resync_block(node, hs[0]) == 0 and bm.node_shard_index(node, hs[0]) == 2
m = bm.metrics()
print("bm smoke ok in %.1f s: reconstruct_calls %d, corruption %d" % (time.time() - t0, m["reconstruct_calls"], m["corruption_counter"]))
bm.close()

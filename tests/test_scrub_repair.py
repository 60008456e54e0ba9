"""BASELINE config 5 in miniature: mixed RS(6,3) / RS(10,4) stripes, every shard independently
corrupted with p = 0.10, one sweep = detect (per-shard blake2sum) -> reconstruct -> rewrite.
Stripes with more than m bad shards must be reported unrecoverable, not crash
(src/block/resync.rs:300-315 error accounting)."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_lib as O  # noqa: E402

import garage_b200 as G  # noqa: E402

pytestmark = pytest.mark.gpu


def make(k, m, n, stride, lens, seed):
    tot = k + m
    data = O.fill_random(n * k * stride, seed)
    P = O.build_matrix(k, m, 0)
    par = O.encode(k, m, P, data, stride, n, lens, simd=True)
    sh = np.zeros((n, tot, stride), dtype=np.uint8)
    sh[:, :k] = data.reshape(n, k, stride)
    sh[:, k:] = par.reshape(n, m, stride)
    for s in range(n):
        sh[s, :, lens[s]:] = 0
    sums = np.zeros((n, tot, 32), dtype=np.uint8)
    for s in range(n):
        for i in range(tot):
            sums[s, i] = np.frombuffer(hashlib.blake2b(sh[s, i, : lens[s]].tobytes()).digest()[:32], dtype=np.uint8)
    return sh, sums


@pytest.mark.parametrize("k,m", [(6, 3), (10, 4)])
@pytest.mark.parametrize("mode", ["device", "host"])
def test_sweep_detects_and_heals(k, m, mode):
    import torch

    tot, stride, n = k + m, 2048 + 128, 300
    rng = np.random.default_rng(k * 7 + (mode == "host"))
    lens = rng.integers(1, stride + 1, n).astype(np.uint32)
    orig, sums = make(k, m, n, stride, lens, 11 + k)
    injected = rng.random((n, tot)) < 0.10
    broken = orig.copy()
    for s, i in zip(*np.nonzero(injected)):
        nflip = int(rng.integers(1, 4))
        for _ in range(nflip):
            broken[s, i, int(rng.integers(0, lens[s]))] ^= int(rng.integers(1, 256))
    # a flip may cancel itself; recompute the truth
    injected = np.array([[not np.array_equal(broken[s, i, : lens[s]], orig[s, i, : lens[s]]) for i in range(tot)]
                         for s in range(n)])
    nbad = injected.sum(axis=1)
    assert (nbad > m).any() and (nbad == 0).any()  # the seed covers both extremes
    with G.GarageEc(0, k, m) as ec:
        if mode == "device":
            d = torch.from_numpy(broken.reshape(-1).copy()).cuda()
            bad = torch.full((n * tot,), 7, dtype=torch.uint8, device="cuda")
            status = torch.full((n,), 7, dtype=torch.int32, device="cuda")
            ec.scrub_repair(d, torch.from_numpy(sums.reshape(-1)).cuda(), bad, stride, n, status=status,
                            shard_len=torch.from_numpy(lens.astype(np.int32)).cuda())
            got, bad_h, st_h = d.cpu().numpy().reshape(n, tot, stride), bad.cpu().numpy(), status.cpu().numpy()
        else:
            buf = broken.reshape(-1).copy()
            bad_h = np.full(n * tot, 7, dtype=np.uint8)
            st_h = np.full(n, 7, dtype=np.int32)
            rc = ec.scrub_repair(buf, sums.reshape(-1), bad_h, stride, n, status=st_h, shard_len=lens)
            assert rc == G.E_UNRECOVERABLE
            got = buf.reshape(n, tot, stride)
        assert np.array_equal(bad_h.reshape(n, tot).astype(bool), injected)
        for s in range(n):
            L = int(lens[s])
            if nbad[s] > m:
                assert st_h[s] == G.E_UNRECOVERABLE
                assert np.array_equal(got[s], broken[s])  # left as found
            else:
                assert st_h[s] == 0
                assert np.array_equal(got[s][:, :L], orig[s][:, :L]), s

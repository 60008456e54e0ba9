"""blake2sum = BLAKE2b-512 truncated to 32 bytes (src/util/data.rs:130-138).  PINNED: python's
hashlib is an independent implementation of RFC 7693, so this component has a real external
oracle.  CPU part: the library's host function; GPU part: the per-shard kernel."""
import hashlib
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_lib as O  # noqa: E402

import garage_b200 as G  # noqa: E402


def ref(b):
    return hashlib.blake2b(bytes(b)).digest()[:32]


def test_rfc7693_abc_vector():
    # RFC 7693 appendix A: BLAKE2b-512("abc") starts with ba80a53f981c4d0d...
    want = bytes.fromhex("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1")
    assert ref(b"abc") == want
    assert G.blake2sum(b"abc") == want


def test_host_blake2sum_matches_hashlib():
    for n in [0, 1, 3, 55, 127, 128, 129, 255, 256, 257, 1000, 4096, 104858, 1 << 20]:
        b = O.fill_random(n, 17 + n)
        assert G.blake2sum(b) == ref(b), n


@pytest.mark.gpu
def test_shard_sums_device_and_host():
    import torch

    k, m = 10, 4
    tot, stride, n = k + m, 4096 + 128, 9
    lens = np.array([stride, 0, 1, 15, 16, 17, 127, 128, 4097], dtype=np.uint32)
    sh = O.fill_random(n * tot * stride, 5).reshape(n, tot, stride)
    want = np.zeros((n, tot, 32), dtype=np.uint8)
    for s in range(n):
        for i in range(tot):
            want[s, i] = np.frombuffer(ref(sh[s, i, : lens[s]]), dtype=np.uint8)
    with G.GarageEc(0, k, m) as ec:
        d = torch.from_numpy(sh.reshape(-1)).cuda()
        sums = torch.zeros(n * tot * 32, dtype=torch.uint8, device="cuda")
        ec.shard_sums(d, sums, stride, n, tot, shard_len=torch.from_numpy(lens.astype(np.int32)).cuda())
        assert np.array_equal(sums.cpu().numpy().reshape(n, tot, 32), want)
        hs = np.zeros(n * tot * 32, dtype=np.uint8)
        ec.shard_sums(sh.reshape(-1), hs, stride, n, tot, shard_len=lens)
        assert np.array_equal(hs.reshape(n, tot, 32), want)
        # check mode: flip a bit in three shards
        bad_sh = sh.copy()
        bad_sh[0, 3, 100] ^= 1
        bad_sh[8, 13, 4096] ^= 0x80
        bad_sh[2, 0, 0] ^= 4
        bad_sh[2, 1, 5] ^= 4  # beyond shard_len (1): not covered, not flagged
        bad = np.full(n * tot, 9, dtype=np.uint8)
        ec.check_sums(bad_sh.reshape(-1), want.reshape(-1), bad, stride, n, tot, shard_len=lens)
        exp = np.zeros((n, tot), dtype=np.uint8)
        exp[0, 3] = exp[8, 13] = exp[2, 0] = 1
        assert np.array_equal(bad.reshape(n, tot), exp)
        bd = torch.zeros(n * tot, dtype=torch.uint8, device="cuda")
        ec.check_sums(torch.from_numpy(bad_sh.reshape(-1)).cuda(), torch.from_numpy(want.reshape(-1)).cuda(), bd,
                      stride, n, tot, shard_len=torch.from_numpy(lens.astype(np.int32)).cuda())
        assert np.array_equal(bd.cpu().numpy().reshape(n, tot), exp)


@pytest.mark.gpu
def test_shard_sums_full_size_sample():
    import torch

    k, m, n = 10, 4, 512
    with G.GarageEc(0, k, m) as ec:
        L = ec.shard_len(1 << 20)
        stride = ec.stride_for(L)
        d = torch.empty(n * k * stride, dtype=torch.uint8, device="cuda")
        ec.fill_random(d, n * k * stride, 3, 0)
        lens = torch.full((n,), L, dtype=torch.int32, device="cuda")
        sums = torch.zeros(n * k * 32, dtype=torch.uint8, device="cuda")
        ec.shard_sums(d, sums, stride, n, k, shard_len=lens)
        h = sums.cpu().numpy().reshape(n * k, 32)
        d3 = d.view(n * k, stride)
        for i in (0, 1, 777, n * k - 1):
            assert h[i].tobytes() == ref(d3[i, :L].cpu().numpy()), i


@pytest.mark.gpu
def test_shard_sums_many_small_shards_scalar_kernel():
    """>= 24 000 shards takes the one-thread-per-shard kernel; fewer takes the quad kernel"""
    import torch

    k, m, n, stride = 10, 4, 2000, 208
    tot = k + m
    rng = np.random.default_rng(3)
    lens = rng.integers(0, stride + 1, n).astype(np.uint32)
    sh = O.fill_random(n * tot * stride, 9).reshape(n, tot, stride)
    with G.GarageEc(0, k, m) as ec:
        d = torch.from_numpy(sh.reshape(-1)).cuda()
        dl = torch.from_numpy(lens.astype(np.int32)).cuda()
        sums = torch.zeros(n * tot * 32, dtype=torch.uint8, device="cuda")
        ec.shard_sums(d, sums, stride, n, tot, shard_len=dl)          # 28 000 shards: scalar kernel
        h = sums.cpu().numpy().reshape(n, tot, 32)
        sums2 = torch.zeros(n * k * 32, dtype=torch.uint8, device="cuda")
        dk = torch.from_numpy(np.ascontiguousarray(sh[:, :k]).reshape(-1)).cuda()
        ec.shard_sums(dk, sums2, stride, n, k, shard_len=dl)          # 20 000 shards: quad kernel
        h2 = sums2.cpu().numpy().reshape(n, k, 32)
        assert np.array_equal(h[:, :k], h2)
        for s in rng.integers(0, n, 40):
            for i in (0, 5, 13):
                assert h[s, i].tobytes() == ref(sh[s, i, : lens[s]]), (s, i)

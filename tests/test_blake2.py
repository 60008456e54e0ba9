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


# ------------------------------------------------------------------ adler8: the fast per-shard tag
# PINNED by python's zlib.adler32 (an independent implementation of RFC 1950's Adler-32): the tag
# is the 8 little-endian Adler-32 values of the shard's 8 segments of roundup16(ceil(len/8)) bytes.
def adler8_ref(b):
    import struct
    import zlib

    b = bytes(b)
    seg = (((len(b) + 7) // 8) + 15) // 16 * 16
    return b"".join(struct.pack("<I", zlib.adler32(b[s * seg:(s + 1) * seg]) & 0xFFFFFFFF) for s in range(8))


def test_adler8_rfc1950_vector_and_host_function():
    import zlib

    assert zlib.adler32(b"Wikipedia") == 0x11E60398  # the published worked example of Adler-32
    assert adler8_ref(b"") == b"\x01\x00\x00\x00" * 8
    for n in [0, 1, 3, 15, 16, 17, 127, 128, 129, 1000, 4096, 5552 * 8, 5553 * 8 + 5, 104858, 174763, 1 << 20]:
        b = O.fill_random(n, 23 + n)
        assert G.shard_sum_host(G.SUM_ADLER8, b) == adler8_ref(b), n
        assert G.shard_sum_host(G.SUM_BLAKE2, b) == ref(b), n
    worst = np.full(1 << 20, 255, dtype=np.uint8)  # largest possible sums: no 32/64-bit overflow anywhere
    assert G.shard_sum_host(G.SUM_ADLER8, worst) == adler8_ref(worst)


def test_adler8_host_vector_path_every_length_class_and_alignment():
    """the host function runs an AVX2 loop (32-byte blocks, runs of at most 5552 bytes, scalar tails) where the CPU
    has it: lengths around every boundary of that loop, unaligned starts, all-0xff data (largest sums) and segments
    spanning many runs, all against zlib"""
    rng = np.random.default_rng(77)
    sizes = [31, 32, 33, 63, 64, 65, 255, 256, 257, 8 * 31, 8 * 32, 8 * 33, 5551, 5552, 5553, 8 * 5552 - 1, 8 * 5552, 8 * 5552 + 9,
             8 * 5536, 8 * 5536 + 32 * 8, 44415, 104857, 262144, (1 << 20) + 13, 3000001]
    for n in sizes:
        for fill in ("rand", "ff"):
            a = rng.integers(0, 256, n, dtype=np.uint8) if fill == "rand" else np.full(n, 255, dtype=np.uint8)
            for off in (0, 1, 7):
                b = np.concatenate([np.zeros(off, dtype=np.uint8), a])[off:]  # start at an odd address
                assert G.shard_sum_host(G.SUM_ADLER8, b) == adler8_ref(a), (n, fill, off)


@pytest.mark.gpu
def test_adler8_device_and_host_paths():
    import torch

    k, m = 10, 4
    tot, stride, n = k + m, 4096 + 128, 10
    lens = np.array([stride, 0, 1, 15, 16, 17, 127, 128, 4097, 2049], dtype=np.uint32)
    sh = O.fill_random(n * tot * stride, 6).reshape(n, tot, stride)
    sh[9] = 255  # worst case sums
    want = np.zeros((n, tot, 32), dtype=np.uint8)
    for s in range(n):
        for i in range(tot):
            want[s, i] = np.frombuffer(adler8_ref(sh[s, i, : lens[s]]), dtype=np.uint8)
    with G.GarageEc(0, k, m) as ec:
        ec.set_sum_kind(G.SUM_ADLER8)
        d = torch.from_numpy(sh.reshape(-1)).cuda()
        dl = torch.from_numpy(lens.astype(np.int32)).cuda()
        sums = torch.zeros(n * tot * 32, dtype=torch.uint8, device="cuda")
        ec.shard_sums(d, sums, stride, n, tot, shard_len=dl)
        assert np.array_equal(sums.cpu().numpy().reshape(n, tot, 32), want)
        hs = np.zeros(n * tot * 32, dtype=np.uint8)
        ec.shard_sums(sh.reshape(-1), hs, stride, n, tot, shard_len=lens)
        assert np.array_equal(hs.reshape(n, tot, 32), want)
        bad_sh = sh.copy()
        bad_sh[0, 3, 100] ^= 1
        bad_sh[8, 13, 4096] ^= 0x80
        bad_sh[2, 0, 0] ^= 4
        bad_sh[2, 1, 5] ^= 4  # beyond shard_len (1): not covered, not flagged
        exp = np.zeros((n, tot), dtype=np.uint8)
        exp[0, 3] = exp[8, 13] = exp[2, 0] = 1
        bd = torch.zeros(n * tot, dtype=torch.uint8, device="cuda")
        ec.check_sums(torch.from_numpy(bad_sh.reshape(-1)).cuda(), torch.from_numpy(want.reshape(-1)).cuda(), bd,
                      stride, n, tot, shard_len=dl)
        assert np.array_equal(bd.cpu().numpy().reshape(n, tot), exp)
        # scrub + repair with the fast tag: the flipped shards are found and rebuilt
        enc_sh = sh.copy()
        P = O.build_matrix(k, m, 0)
        for s in range(n):
            enc_sh[s, k:] = O.encode(k, m, P, np.ascontiguousarray(enc_sh[s, :k]).reshape(-1), stride, 1,
                                     lens[s:s + 1]).reshape(m, stride)
        tags = np.zeros(n * tot * 32, dtype=np.uint8)
        ec.shard_sums(enc_sh.reshape(-1), tags, stride, n, tot, shard_len=lens)
        hurt = enc_sh.copy()
        hurt[0, 2, 7] ^= 1
        hurt[8, 12, 4000] ^= 2
        bad = np.zeros(n * tot, dtype=np.uint8)
        st = np.zeros(n, dtype=np.int32)
        ec.scrub_repair(hurt.reshape(-1), tags, bad, stride, n, status=st, shard_len=lens)
        assert bad.sum() == 2 and not st.any()
        for s in range(n):
            assert np.array_equal(hurt[s, :, : lens[s]], enc_sh[s, :, : lens[s]]), s


@pytest.mark.gpu
def test_adler8_full_size_shards_and_block_level_sums():
    import torch

    k, m, n = 10, 4, 64
    blocks = [O.fill_random((1 << 20) - 11 * i, 900 + i) for i in range(n)]
    with G.GarageEc(0, k, m) as ec:
        ec.set_sum_kind(G.SUM_ADLER8)
        stride = ec.stride_for(ec.shard_len(1 << 20))
        par = np.zeros(n * m * stride, dtype=np.uint8)
        sums = np.zeros(n * (k + m) * 32, dtype=np.uint8)
        ec.encode_blocks(blocks, par, stride, sums_out=sums)
        sums = sums.reshape(n, k + m, 32)
        for s in (0, 1, 31, n - 1):
            L = ec.shard_len(blocks[s].size)
            d = O.split_block(blocks[s], k, stride).reshape(k, stride)
            for j in (0, k - 1):
                assert sums[s, j].tobytes() == adler8_ref(d[j, :L]), (s, j)
            for i in (0, m - 1):
                assert sums[s, k + i].tobytes() == adler8_ref(par.reshape(n, m, stride)[s, i, :L]), (s, i)

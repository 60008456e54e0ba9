"""CPU-only tests: pin the oracle (oracle/rs_oracle.c) against the survey-time known-answer
vectors (tests/golden/kat.json), against the independent numpy
restatement (oracle/rs_oracle_np.py), and by algebraic properties (field axioms, MDS,
encode->erase->decode round trip).  The reference itself has no RS code (parity unpinned)."""
import itertools
import json
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as O  # noqa: E402

sys.path.insert(0, os.path.join(O.ROOT, "oracle"))
import rs_oracle_np as NP  # noqa: E402

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
SEED = 0x6761726167650010


def hx(s):
    return np.array([int(x, 16) for x in s.replace("/", " ").replace("|", " ").split()], dtype=np.uint8)


# ---------------------------------------------------------------- field
def test_field_kat():
    L = O.lib()
    assert [L.rs_oracle_gf_exp(i) for i in range(16)] == list(hx(KAT["exp_0_16"]))
    assert [L.rs_oracle_gf_log(i) for i in range(1, 9)] == KAT["log_1_8"]
    for a, b, c in KAT["mul"]:
        assert L.rs_oracle_gf_mul(int(a, 16), int(b, 16)) == int(c, 16)
        assert NP.gf_mul_scalar(int(a, 16), int(b, 16)) == int(c, 16)
    for a, b in KAT["inv"]:
        assert L.rs_oracle_gf_inv(int(a, 16)) == int(b, 16)
        assert NP.gf_inv_scalar(int(a, 16)) == int(b, 16)


def test_field_axioms_and_cross_impl():
    L = O.lib()
    tab = np.array([[L.rs_oracle_gf_mul(a, b) for b in range(256)] for a in range(256)], dtype=np.uint8)
    assert np.array_equal(tab, NP.MUL)  # log/exp vs shift-xor: full 64K table
    assert np.array_equal(tab, tab.T)
    assert np.all(tab[1] == np.arange(256))
    assert np.all(tab[0] == 0)
    for a in range(1, 256):
        assert tab[a, L.rs_oracle_gf_inv(a)] == 1
    rng = np.random.default_rng(1)
    a, b, c = rng.integers(0, 256, (3, 2000))
    assert np.array_equal(tab[a, b ^ c], tab[a, b] ^ tab[a, c])  # distributive
    assert np.array_equal(tab[tab[a, b], c], tab[a, tab[b, c]])  # associative


# ---------------------------------------------------------------- matrices
@pytest.mark.parametrize("kind,name", [(0, "vandermonde"), (1, "cauchy")])
@pytest.mark.parametrize("km", ["4,2", "6,3", "10,4"])
def test_matrix_kat(kind, name, km):
    k, m = map(int, km.split(","))
    want = hx(KAT["matrices"][name][km]).reshape(m, k)
    assert np.array_equal(O.build_matrix(k, m, kind), want)
    assert np.array_equal(NP.build_matrix(k, m, kind), want)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", [(1, 1), (2, 1), (3, 2), (5, 3), (8, 4), (12, 4), (16, 4), (20, 8), (32, 8)])
def test_matrix_cross_impl(k, m, kind):
    assert np.array_equal(O.build_matrix(k, m, kind), NP.build_matrix(k, m, kind))


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", [(4, 2), (6, 3), (10, 4)])
def test_mds(k, m, kind):
    """every k-row subset of [I;P] is invertible"""
    P = O.build_matrix(k, m, kind)
    G = NP.generator(k, m, P)
    L = O.lib()
    for rows in itertools.combinations(range(k + m), k):
        sub = np.ascontiguousarray(G[list(rows)])
        assert L.rs_oracle_invert(sub.ctypes.data, k) == 0, rows


def test_invert_roundtrip_and_singular():
    rng = np.random.default_rng(2)
    L = O.lib()
    for n in (1, 2, 5, 10, 17):
        for _ in range(10):
            A = rng.integers(0, 256, (n, n), dtype=np.uint8)
            Ai = A.copy()
            if L.rs_oracle_invert(Ai.ctypes.data, n):
                continue
            assert np.array_equal(NP.gf_matmul(A, Ai), np.eye(n, dtype=np.uint8))
    Z = np.array([[1, 2], [1, 2]], dtype=np.uint8)
    assert L.rs_oracle_invert(Z.ctypes.data, 2) == -1


# ---------------------------------------------------------------- encode
def test_tiny_kat():
    t = KAT["tiny"]
    k, m, L = t["k"], t["m"], t["shard_len"]
    data = hx(t["data"])
    for kind, key in ((0, "vandermonde_parity"), (1, "cauchy_parity")):
        P = O.build_matrix(k, m, kind)
        want = hx(t[key])
        assert np.array_equal(O.encode(k, m, P, data, L, 1), want)
        assert np.array_equal(O.encode(k, m, P, data, L, 1, simd=True), want)
        assert np.array_equal(NP.encode(P, data.reshape(k, L)).reshape(-1), want)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", [(4, 2), (6, 3), (10, 4), (3, 1), (12, 5), (17, 8)])
def test_encode_cross_impl(k, m, kind):
    stride, n = 160, 3
    lens = np.array([160, 97, 1], dtype=np.uint32)
    data = O.fill_random(n * k * stride, SEED + k)
    P = O.build_matrix(k, m, kind)
    par = O.encode(k, m, P, data, stride, n, lens)
    d3 = data.reshape(n, k, stride)
    p3 = par.reshape(n, m, stride)
    for s in range(n):
        L = int(lens[s])
        assert np.array_equal(NP.encode(P, d3[s][:, :L]), p3[s][:, :L])
        assert not p3[s][:, L:].any()  # oracle never writes past shard_len


def test_linearity():
    k, m, L = 10, 4, 257
    P = O.build_matrix(k, m, 0)
    a = O.fill_random(k * L, 1)
    b = O.fill_random(k * L, 2)
    pa, pb, pab = (O.encode(k, m, P, x, L, 1) for x in (a, b, a ^ b))
    assert np.array_equal(pa ^ pb, pab)


# ---------------------------------------------------------------- SIMD CPU baseline == oracle
@pytest.mark.parametrize("isa", [0, 1, 2])
def test_simd_matches_oracle_all_isa(isa):
    L = O.lib()
    got = L.rs_simd_force_isa(isa)
    try:
        if got != isa:
            pytest.skip("ISA %d not available on this CPU" % isa)
        for k, m in ((4, 2), (6, 3), (10, 4), (13, 6)):
            stride, n = 4096 + 192, 5
            lens = np.array([stride, 4096 + 65, 63, 1, 4097], dtype=np.uint32)
            data = O.fill_random(n * k * stride, SEED ^ k)
            P = O.build_matrix(k, m, 0)
            want = O.encode(k, m, P, data, stride, n, lens)
            for th in (1, 3):
                assert np.array_equal(O.encode(k, m, P, data, stride, n, lens, simd=True, threads=th), want)
            # reconstruct + verify
            tot = k + m
            shards = np.zeros((n, tot, stride), dtype=np.uint8)
            shards[:, :k] = data.reshape(n, k, stride)
            shards[:, k:] = want.reshape(n, m, stride)
            for s in range(n):
                shards[s, :, lens[s]:] = 0
            orig = shards.copy()
            rng = np.random.default_rng(isa * 100 + k)
            present = np.ones((n, tot), dtype=np.uint8)
            for s in range(n):
                present[s, rng.choice(tot, size=rng.integers(0, m + 1), replace=False)] = 0
                shards[s, present[s] == 0] = 0xEE
            a = shards.copy()
            b = shards.copy()
            bad_a, st_a = O.reconstruct(k, m, P, a.reshape(-1), present, stride, n, lens)
            bad_b, st_b = O.reconstruct(k, m, P, b.reshape(-1), present, stride, n, lens, simd=True, threads=2)
            assert bad_a == bad_b == 0 and np.array_equal(st_a, st_b)
            for s in range(n):
                assert np.array_equal(a[s, :, : lens[s]], orig[s, :, : lens[s]])
                assert np.array_equal(b[s, :, : lens[s]], orig[s, :, : lens[s]])
            orig2 = orig.copy()
            orig2[1, k + 1, 5] ^= 1
            orig2[2, 0, 0] ^= 0x80
            mm_a = O.verify(k, m, P, orig2.reshape(-1), stride, n, lens)
            mm_b = O.verify(k, m, P, orig2.reshape(-1), stride, n, lens, simd=True, threads=2)
            assert np.array_equal(mm_a, mm_b)
            assert mm_a[0] == 0 and mm_a[1] == 2 and mm_a[2] != 0
    finally:
        L.rs_simd_force_isa(-1)


# ---------------------------------------------------------------- decode
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", [(4, 2), (6, 3), (10, 4)])
def test_reconstruct_all_patterns_small(k, m, kind):
    """every erasure pattern of weight <= m, oracle (2-step) vs numpy (composed matrix)"""
    L = 33
    tot = k + m
    P = O.build_matrix(k, m, kind)
    data = O.fill_random(k * L, 7 * k + kind).reshape(k, L)
    full = np.concatenate([data, NP.encode(P, data)], axis=0)
    pats = [c for e in range(0, m + 1) for c in itertools.combinations(range(tot), e)]
    if len(pats) > 400:
        rng = np.random.default_rng(5)
        pats = [pats[i] for i in rng.choice(len(pats), 400, replace=False)]
    n = len(pats)
    shards = np.tile(full[None], (n, 1, 1)).copy()
    present = np.ones((n, tot), dtype=np.uint8)
    for s, pat in enumerate(pats):
        present[s, list(pat)] = 0
        shards[s, list(pat)] = 0x5A
    npv = shards.copy()
    bad, status = O.reconstruct(k, m, P, shards.reshape(-1), present, L, n)
    assert bad == 0 and not status.any()
    assert np.array_equal(shards, np.tile(full[None], (n, 1, 1)))
    for s in range(0, n, 7):
        assert NP.reconstruct(k, m, P, npv[s], present[s])
        assert np.array_equal(npv[s], full)


def test_unrecoverable_reported_not_fatal():
    k, m, L, n = 6, 3, 16, 4
    tot = k + m
    P = O.build_matrix(k, m, 0)
    data = O.fill_random(n * k * L, 3)
    par = O.encode(k, m, P, data, L, n)
    shards = np.concatenate([data.reshape(n, k, L), par.reshape(n, m, L)], axis=1).copy()
    orig = shards.copy()
    present = np.ones((n, tot), dtype=np.uint8)
    present[1, [0, 2, 4, 8]] = 0  # 4 > m missing
    present[3, [1]] = 0
    shards[1, [0, 2, 4, 8]] = 0
    shards[3, 1] = 0
    snap = shards.copy()
    bad, status = O.reconstruct(k, m, P, shards.reshape(-1), present, L, n)
    assert bad == 1 and list(status) == [0, -1, 0, 0]
    assert np.array_equal(shards[1], snap[1])  # untouched
    assert np.array_equal(shards[3], orig[3])
    assert NP.decode_matrix(k, m, P, present[1]) is None


# ---------------------------------------------------------------- framing + generator
@pytest.mark.parametrize("k", [4, 6, 10])
@pytest.mark.parametrize("blen", [1, 3, 3072, 3073, 1048575, 1048576])
def test_split_join(k, blen):
    L = O.lib().rs_oracle_shard_len(blen, k)
    assert L == NP.shard_len(blen, k) == -(-blen // k)
    stride = (L + 127) // 128 * 128
    block = O.fill_random(blen, blen)
    sh = O.split_block(block, k, stride)
    ref = NP.split_block(block, k)
    assert np.array_equal(sh.reshape(k, stride)[:, :L], ref)
    assert np.array_equal(O.join_block(sh, blen, k, stride), block)


def test_shard_len_configs():
    # SURVEY.md section 7 "Shard geometry": 262144 (k=4), 174763 (k=6), 104858 (k=10)
    L = O.lib()
    assert L.rs_oracle_shard_len(1 << 20, 4) == 262144
    assert L.rs_oracle_shard_len(1 << 20, 6) == 174763
    assert L.rs_oracle_shard_len(1 << 20, 10) == 104858


def test_fill_random_stream():
    a = O.fill_random(1000, SEED, 0)
    assert np.array_equal(a, NP.fill_random(1000, SEED, 0))
    b = O.fill_random(200, SEED, 800)
    assert np.array_equal(b, a[800:1000])
    assert np.array_equal(NP.fill_random(64, SEED, 512), a[512:576])
    # not degenerate
    assert len(set(a.tolist())) > 200

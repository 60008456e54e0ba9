"""GPU parity tests: every call goes through the C ABI (libgarage_ec.so via ctypes) and is
compared bit for bit with the CPU oracle on the same seeded inputs, with the committed golden
digests, and -- at BASELINE.json's full sizes -- through size-independent properties
(encode -> erase -> reconstruct round trip, verify == 0, linearity, digest of digests)."""
import hashlib
import itertools
import json
import os
import sys
import threading

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as O  # noqa: E402
import make_golden as MG  # noqa: E402

import garage_b200 as G  # noqa: E402

pytestmark = pytest.mark.gpu
SEED = 0x6761726167650010
GOLD = json.load(open(os.path.join(HERE, "golden", "parity_digests.json")))["cases"]


@pytest.fixture(scope="module")
def torch():
    import torch as t

    assert t.cuda.is_available()
    return t


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


CODES = [(4, 2), (6, 3), (10, 4), (3, 1), (1, 1), (5, 4), (12, 5), (13, 4), (17, 8), (32, 8), (2, 8)]


# ------------------------------------------------------------------ matrix + boundary
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", CODES)
def test_matrix_matches_oracle(torch, k, m, kind):
    with G.GarageEc(0, k, m, kind) as ec:
        assert np.array_equal(ec.matrix(), O.build_matrix(k, m, kind))


def test_create_with_matrix_roundtrip(torch):
    P = O.build_matrix(10, 4, 1)
    with G.GarageEc(0, 10, 4, matrix=P) as ec:
        assert np.array_equal(ec.matrix(), P)


def test_alignment_and_geometry_errors(torch):
    with G.GarageEc(0, 4, 2) as ec:
        buf = torch.zeros(4 * 64 + 2 * 64 + 64, dtype=torch.uint8, device="cuda")
        with pytest.raises(G.EcError) as e:
            ec.encode(buf[1:], buf[4 * 64:], 64, 1)
        assert e.value.code == G.E_ALIGN
        with pytest.raises(G.EcError) as e:
            ec.encode(buf, buf[4 * 64:], 60, 1)
        assert e.value.code == G.E_ALIGN
        assert ec.encode(buf, buf[4 * 64:], 64, 0) == 0  # empty batch is fine


# ------------------------------------------------------------------ encode
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", CODES)
def test_encode_ragged_device(torch, k, m, kind):
    stride = 1200 * 16
    lens = np.array([stride, 1, 15, 16, 17, 511, 512, 513, 4097, stride - 1, stride - 16, 8191],
                    dtype=np.uint32)
    n = len(lens)
    data = O.fill_random(n * k * stride, SEED + 31 * k + m)
    P = O.build_matrix(k, m, kind)
    want = O.encode(k, m, P, data, stride, n, lens)
    with G.GarageEc(0, k, m, kind) as ec:
        d = dev(torch, data)
        par = torch.full((n * m * stride,), 0xAB, dtype=torch.uint8, device="cuda")
        ec.encode(d, par, stride, n, shard_len=dev(torch, lens))
        got = host(par).reshape(n, m, stride)
        w3 = want.reshape(n, m, stride)
        for s in range(n):
            L = int(lens[s])
            L16 = (L + 15) // 16 * 16
            assert np.array_equal(got[s][:, :L], w3[s][:, :L]), (s, L)
            assert not got[s][:, L:L16].any()           # pad of the last vector is zeroed
            assert (got[s][:, L16:] == 0xAB).all()      # nothing written beyond it
        # uniform length path (shard_len = NULL)
        par2 = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
        ec.encode(d, par2, stride, n)
        assert np.array_equal(host(par2), O.encode(k, m, P, data, stride, n))


def test_encode_ignores_garbage_in_input_pad(torch):
    k, m, stride, L = 10, 4, 256, 100
    data = O.fill_random(k * stride, 5).reshape(k, stride)
    clean = data.copy()
    clean[:, L:] = 0
    P = O.build_matrix(k, m, 0)
    want = O.encode(k, m, P, clean.reshape(-1), stride, 1, np.array([L], dtype=np.uint32))
    with G.GarageEc(0, k, m) as ec:
        par = torch.zeros(m * stride, dtype=torch.uint8, device="cuda")
        ec.encode(dev(torch, data.reshape(-1)), par, stride, 1, shard_len=dev(torch, np.array([L], dtype=np.uint32)))
        assert np.array_equal(host(par), want)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_encode_golden(torch, name):
    g = GOLD[name]
    k, m, kind, n = g["k"], g["m"], g["kind"], g["n"]
    bl = g["block_lens"][0] if len(set(g["block_lens"])) == 1 else None
    data, L, stride, _ = MG.make_case(k, m, kind, n, bl)
    with G.GarageEc(0, k, m, kind) as ec:
        par = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
        ec.encode(dev(torch, data), par, stride, n, shard_len=dev(torch, L))
        h = host(par)
        assert h[:64].tobytes().hex() == g["parity_first64"]
        assert hashlib.sha256(h.tobytes()).hexdigest() == g["parity_sha256"]
        # HOST mode through the staging lanes gives the same bytes
        ph = np.zeros(n * m * stride, dtype=np.uint8)
        ec.encode(data, ph, stride, n, shard_len=L)
        assert hashlib.sha256(ph.tobytes()).hexdigest() == g["parity_sha256"]


def test_linearity_device(torch):
    k, m, stride, n = 10, 4, 104960, 3
    a = O.fill_random(n * k * stride, 11)
    b = O.fill_random(n * k * stride, 12)
    with G.GarageEc(0, k, m) as ec:
        outs = []
        for x in (a, b, a ^ b):
            p = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
            ec.encode(dev(torch, x), p, stride, n)
            outs.append(p)
        assert torch.equal(outs[0] ^ outs[1], outs[2])


# ------------------------------------------------------------------ reconstruct
def make_stripes(k, m, kind, n, stride, lens, seed):
    tot = k + m
    data = O.fill_random(n * k * stride, seed)
    P = O.build_matrix(k, m, kind)
    par = O.encode(k, m, P, data, stride, n, lens, simd=True)
    sh = np.zeros((n, tot, stride), dtype=np.uint8)
    sh[:, :k] = data.reshape(n, k, stride)
    sh[:, k:] = par.reshape(n, m, stride)
    if lens is not None:
        for s in range(n):
            sh[s, :, lens[s]:] = 0
    return P, sh


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("k,m", [(4, 2), (6, 3), (10, 4)])
def test_reconstruct_every_pattern(torch, k, m, kind):
    tot, stride = k + m, 272
    pats = [c for e in range(0, m + 1) for c in itertools.combinations(range(tot), e)]
    n = len(pats)
    lens = np.array([1 + (37 * i) % stride for i in range(n)], dtype=np.uint32)
    P, orig = make_stripes(k, m, kind, n, stride, lens, 77 + k)
    present = np.ones((n, tot), dtype=np.uint8)
    broken = orig.copy()
    for s, pat in enumerate(pats):
        present[s, list(pat)] = 0
        broken[s, list(pat)] = 0xC3
    ref = broken.copy()
    bad, st_ref = O.reconstruct(k, m, P, ref.reshape(-1), present, stride, n, lens)
    assert bad == 0
    with G.GarageEc(0, k, m, kind) as ec:
        d = dev(torch, broken.reshape(-1))
        status = torch.full((n,), 99, dtype=torch.int32, device="cuda")
        ec.reconstruct(d, dev(torch, present), stride, n, status=status, shard_len=dev(torch, lens))
        got = host(d).reshape(n, tot, stride)
        assert not host(status).any()
        for s in range(n):
            L = int(lens[s])
            assert np.array_equal(got[s][:, :L], orig[s][:, :L]), (s, pats[s])
            assert np.array_equal(got[s][:, :L], ref[s][:, :L])
            # present shards are never written
            pres = present[s].astype(bool)
            assert np.array_equal(got[s][pres], broken[s][pres])


@pytest.mark.parametrize("k,m", [(12, 5), (17, 8), (32, 8), (3, 1), (2, 8)])
def test_reconstruct_generic_k_and_two_pass_m(torch, k, m):
    tot, stride, n = k + m, 4096 + 64, 24
    rng = np.random.default_rng(k * 100 + m)
    lens = rng.integers(1, stride + 1, n).astype(np.uint32)
    P, orig = make_stripes(k, m, 0, n, stride, lens, 5 + k)
    present = np.ones((n, tot), dtype=np.uint8)
    broken = orig.copy()
    for s in range(n):
        e = int(rng.integers(0, m + 1)) if s else m
        idx = rng.choice(tot, e, replace=False)
        present[s, idx] = 0
        broken[s, idx] = 0x11
    with G.GarageEc(0, k, m) as ec:
        d = dev(torch, broken.reshape(-1))
        ec.reconstruct(d, dev(torch, present), stride, n, shard_len=dev(torch, lens))
        got = host(d).reshape(n, tot, stride)
        for s in range(n):
            assert np.array_equal(got[s][:, : lens[s]], orig[s][:, : lens[s]]), s


def test_reconstruct_want_mask_and_unrecoverable(torch):
    k, m, stride, n = 10, 4, 1024, 6
    tot = k + m
    P, orig = make_stripes(k, m, 0, n, stride, None, 9)
    present = np.ones((n, tot), dtype=np.uint8)
    want = np.zeros((n, tot), dtype=np.uint8)
    broken = orig.copy()
    present[0, [1, 12]] = 0; want[0, 1] = 1            # GET: data shard only
    present[1, [0, 3, 11, 13]] = 0; want[1, 13] = 1    # resync: my own (parity) shard only
    present[2, [0, 1, 2, 3, 4]] = 0; want[2, :] = 1    # 5 > m: unrecoverable
    present[3, [5]] = 0                                # absent but not wanted: untouched
    present[4, :] = 1; want[4, :] = 1                  # nothing absent
    present[5, [9, 10, 11, 12]] = 0; want[5, :k] = 1   # GET with 4 erasures
    for s in range(n):
        broken[s, present[s] == 0] = 0x77
    with G.GarageEc(0, k, m) as ec:
        d = dev(torch, broken.reshape(-1))
        status = torch.zeros(n, dtype=torch.int32, device="cuda")
        rc = ec.reconstruct(d, dev(torch, present), stride, n, want=dev(torch, want), status=status)
        assert rc == 0  # DEVICE calls only fill status[]
        got = host(d).reshape(n, tot, stride)
        assert list(host(status)) == [0, 0, G.E_UNRECOVERABLE, 0, 0, 0]
        exp = broken.copy()
        exp[0, 1] = orig[0, 1]
        exp[1, 13] = orig[1, 13]
        exp[5, 9] = orig[5, 9]
        assert np.array_equal(got, exp)
        # HOST mode reports it in the return code as well
        hb = broken.copy().reshape(-1)
        hs = np.zeros(n, dtype=np.int32)
        rc = ec.reconstruct(hb, present, stride, n, want=want, status=hs)
        assert rc == G.E_UNRECOVERABLE and list(hs) == [0, 0, G.E_UNRECOVERABLE, 0, 0, 0]
        assert np.array_equal(hb.reshape(n, tot, stride), exp)


def test_reconstruct_shared_pattern_reuses_tables(torch):
    """a whole node lost: every stripe misses the same shard (the resync case)"""
    k, m, stride, n = 10, 4, 8192, 300
    tot = k + m
    P, orig = make_stripes(k, m, 0, n, stride, None, 21)
    present = np.ones((n, tot), dtype=np.uint8)
    present[:, 3] = 0
    present[::7, 12] = 0
    broken = orig.copy()
    broken[:, 3] = 0
    broken[::7, 12] = 0
    with G.GarageEc(0, k, m) as ec:
        d = dev(torch, broken.reshape(-1))
        ec.reconstruct(d, dev(torch, present), stride, n)
        assert np.array_equal(host(d).reshape(n, tot, stride), orig)


# ------------------------------------------------------------------ verify (scrub)
@pytest.mark.parametrize("k,m", [(4, 2), (6, 3), (10, 4), (17, 8)])
def test_verify_flags_exact_rows(torch, k, m):
    tot, stride, n = k + m, 2048, 40
    rng = np.random.default_rng(k)
    lens = rng.integers(1, stride + 1, n).astype(np.uint32)
    P, sh = make_stripes(k, m, 0, n, stride, lens, 3 * k)
    for s in range(0, n, 3):
        i = int(rng.integers(0, tot))
        t = int(rng.integers(0, lens[s]))
        sh[s, i, t] ^= 1 << int(rng.integers(0, 8))
    # garbage past shard_len must not be flagged
    for s in range(n):
        sh[s, :, lens[s]:] = 0xFF
    want = O.verify(k, m, P, sh.reshape(-1), stride, n, lens)
    with G.GarageEc(0, k, m) as ec:
        mm = torch.full((n,), 0xFFFF, dtype=torch.int32, device="cuda")
        ec.verify(dev(torch, sh.reshape(-1)), mm, stride, n, shard_len=dev(torch, lens))
        assert np.array_equal(host(mm).astype(np.uint32), want)
        assert want[1] == 0 and want[0] != 0
        hm = np.zeros(n, dtype=np.uint32)
        ec.verify(sh.reshape(-1), hm, stride, n, shard_len=lens)
        assert np.array_equal(hm, want)


# ------------------------------------------------------------------ block-level host API
@pytest.mark.parametrize("k,m", [(4, 2), (10, 4), (6, 3)])
def test_encode_decode_blocks_host(torch, k, m):
    tot = k + m
    blens = [1 << 20, 3073, 1, 1048575, 65536, 777777, (1 << 20) - 7, 5]
    blocks = [O.fill_random(b, 1000 + i) for i, b in enumerate(blens)]
    n = len(blocks)
    with G.GarageEc(0, k, m) as ec:
        stride = ec.stride_for(ec.shard_len(max(blens)))
        par = np.zeros(n * m * stride, dtype=np.uint8)
        ec.encode_blocks(blocks, par, stride)
        P = O.build_matrix(k, m, 0)
        shards = np.zeros((n, tot, stride), dtype=np.uint8)
        for s, b in enumerate(blocks):
            L = ec.shard_len(b.size)
            d = O.split_block(b, k, stride)
            shards[s, :k] = d.reshape(k, stride)
            want = O.encode(k, m, P, d, stride, 1, np.array([L], dtype=np.uint32)).reshape(m, stride)
            got = par.reshape(n, m, stride)[s]
            assert np.array_equal(got[:, :L], want[:, :L]), s
            shards[s, k:] = got
        # GET with erasures
        rng = np.random.default_rng(k)
        present = np.ones((n, tot), dtype=np.uint8)
        for s in range(n):
            idx = rng.choice(tot, int(rng.integers(0, m + 1)), replace=False)
            present[s, idx] = 0
            shards[s, idx] = 0xEE
        present[2, :] = 1
        present[2, : m + 1] = 0  # unrecoverable
        out = [np.zeros(b, dtype=np.uint8) for b in blens]
        status = np.zeros(n, dtype=np.int32)
        rc = ec.decode_blocks(shards.reshape(-1), present, blens, stride, out, status)
        assert rc == G.E_UNRECOVERABLE and status[2] == G.E_UNRECOVERABLE
        for s in range(n):
            if s != 2:
                assert status[s] == 0 and np.array_equal(out[s], blocks[s]), s


# ------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize("k,m,n", [(10, 4, 4096), (6, 3, 1024), (4, 2, 1024)])
def test_full_size_roundtrip(torch, k, m, n):
    """BASELINE configs 2+3: encode n x 1 MiB blocks, erase m random shards per stripe,
    reconstruct, compare with the originals (every byte); scrub says clean; every parity byte of
    every stripe == CPU oracle."""
    B = 1 << 20
    tot = k + m
    with G.GarageEc(0, k, m) as ec:
        L = ec.shard_len(B)
        stride = ec.stride_for(L)
        shards = torch.zeros(n * tot * stride, dtype=torch.uint8, device="cuda")
        sh3 = shards.view(n, tot, stride)
        # data shards = the block split; generated on device from the shared counter stream
        tmp = torch.empty(n * k * stride, dtype=torch.uint8, device="cuda")
        ec.fill_random(tmp, n * k * stride, SEED, 0)
        t3 = tmp.view(n, k, stride)
        t3[:, :, L:] = 0
        if k * L > B:  # zero-pad the tail of the last data shard like the framing does
            t3[:, k - 1, L - (k * L - B):] = 0
        sh3[:, :k] = t3
        lens = torch.full((n,), L, dtype=torch.int32, device="cuda")
        par = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
        ec.encode(tmp, par, stride, n, shard_len=lens)
        sh3[:, k:] = par.view(n, m, stride)
        torch.cuda.synchronize()
        # EVERY parity byte of EVERY stripe against the CPU oracle: the multi-threaded
        # SIMD arm computes all n stripes, and is itself held to the normative scalar oracle on a sample
        P = O.build_matrix(k, m, 0)
        h_data = host(tmp)
        h_lens = np.full(n, L, dtype=np.uint32)
        want_all = O.encode(k, m, P, h_data, stride, n, h_lens, simd=True)
        got_all = host(par)
        assert got_all.shape == want_all.shape
        if not np.array_equal(got_all, want_all):
            badrow = np.flatnonzero((got_all.reshape(n, -1) != want_all.reshape(n, -1)).any(axis=1))
            raise AssertionError("parity differs from the oracle in %d stripes, first %s" % (badrow.size, badrow[:8]))
        for s in list(range(0, n, max(1, n // 13))) + [n - 1]:
            d = h_data.reshape(n, k * stride)[s]
            scalar = O.encode(k, m, P, d, stride, 1, np.array([L], dtype=np.uint32))
            assert np.array_equal(want_all.reshape(n, m * stride)[s], scalar), ("simd arm vs scalar oracle", s)
        del want_all, got_all, h_data
        # the device data is the documented stream (spot check one stripe against the CPU generator)
        assert np.array_equal(host(t3[3, 0, :L]), O.fill_random(L, SEED, (3 * k) * stride)[:L])
        # scrub: clean
        mm = torch.ones(n, dtype=torch.int32, device="cuda")
        ec.verify(shards, mm, stride, n, shard_len=lens)
        assert int(mm.abs().sum()) == 0
        # erase m random shards per stripe, reconstruct, compare
        orig = shards.clone()
        g = torch.Generator(device="cpu").manual_seed(1234)
        keys = torch.rand(n, tot, generator=g)
        erased = keys.argsort(dim=1)[:, :m]
        present = torch.ones(n, tot, dtype=torch.uint8)
        present.scatter_(1, erased, 0)
        pd = present.cuda()
        sh3[~pd.bool()] = 0
        assert not torch.equal(shards, orig)
        status = torch.ones(n, dtype=torch.int32, device="cuda")
        ec.reconstruct(shards, pd, stride, n, status=status, shard_len=lens)
        assert int(status.abs().sum()) == 0
        assert torch.equal(shards, orig)
        # corrupt 1 byte in a few parity shards -> scrub flags exactly those rows
        for s, i in ((0, 0), (n // 2, m - 1), (n - 1, 0)):
            sh3[s, k + i, L - 1] ^= 0x40
        ec.verify(shards, mm, stride, n, shard_len=lens)
        mmh = host(mm)
        assert mmh[0] == 1 and mmh[n // 2] == 1 << (m - 1) and mmh[n - 1] == 1
        assert int(np.count_nonzero(mmh)) == 3


def test_thread_safety_device_calls(torch):
    k, m, stride, n = 10, 4, 16384, 64
    data = O.fill_random(n * k * stride, 42)
    want = O.encode(k, m, O.build_matrix(k, m, 0), data, stride, n, simd=True)
    with G.GarageEc(0, k, m) as ec:
        d = dev(torch, data)
        errs = []

        def work(i):
            try:
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    p = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
                    for _ in range(5):
                        ec.encode(d, p, stride, n)
                    s.synchronize()
                    if not np.array_equal(host(p), want):
                        errs.append(i)
            except Exception as ex:  # noqa: BLE001
                errs.append(repr(ex))

        th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs


def test_host_mode_concurrent_callers_one_context(torch):
    """HOST-mode calls from several OS threads on ONE context (they share the staging lanes and
    are serialised inside the library) must each get the right bytes."""
    k, m, stride, n = 6, 3, 8192, 40
    P = O.build_matrix(k, m, 0)
    datas = [O.fill_random(n * k * stride, 900 + i) for i in range(4)]
    wants = [O.encode(k, m, P, d, stride, n, simd=True) for d in datas]
    with G.GarageEc(0, k, m) as ec:
        errs = []

        def work(i):
            try:
                for _ in range(3):
                    out = np.zeros(n * m * stride, dtype=np.uint8)
                    ec.encode(datas[i], out, stride, n)
                    if not np.array_equal(out, wants[i]):
                        errs.append(i)
                    sh = np.concatenate([datas[i].reshape(n, k, stride), wants[i].reshape(n, m, stride)], axis=1)
                    mm = np.ones(n, dtype=np.uint32)
                    ec.verify(np.ascontiguousarray(sh).reshape(-1), mm, stride, n)
                    if mm.any():
                        errs.append(("verify", i))
            except Exception as ex:  # noqa: BLE001
                errs.append(repr(ex))

        th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs


def test_device_calls_are_cuda_graph_capturable(torch):
    """DEVICE-mode encode / reconstruct / verify enqueue only kernels and stream-ordered
    allocations, so a caller can capture them once and replay (no host work per batch)."""
    k, m, stride, n = 10, 4, 4096, 64
    tot = k + m
    P, orig = make_stripes(k, m, 0, n, stride, None, 77)
    present = np.ones((n, tot), dtype=np.uint8)
    rng = np.random.default_rng(1)
    for s in range(n):
        present[s, rng.choice(tot, m, replace=False)] = 0
    with G.GarageEc(0, k, m) as ec:
        data = dev(torch, orig[:, :k].reshape(-1))
        par = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
        sh = dev(torch, orig.reshape(-1))
        pd = dev(torch, present)
        st = torch.ones(n, dtype=torch.int32, device="cuda")
        mm = torch.ones(n, dtype=torch.int32, device="cuda")

        def step():
            ec.encode(data, par, stride, n)
            ec.reconstruct(sh, pd, stride, n, status=st)
            ec.verify(sh, mm, stride, n)

        step()  # warm-up outside capture (function attributes, allocator pool)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for trial in range(3):
            par.zero_()
            sh.view(n, tot, stride)[~pd.bool()] = trial  # wipe the erased shards
            st.fill_(9)
            mm.fill_(9)
            g.replay()
            torch.cuda.synchronize()
            assert np.array_equal(host(par).reshape(n, m, stride), orig[:, k:])
            assert np.array_equal(host(sh).reshape(n, tot, stride), orig)
            assert not host(st).any() and not host(mm).any()


# ------------------------------------------------------------------ every launch-shape class of k
@pytest.mark.parametrize("k,m", [(5, 2), (7, 3), (9, 3), (11, 4), (14, 4), (15, 3), (16, 4), (19, 3), (20, 4), (21, 2),
                                 (24, 3), (25, 2), (28, 4), (29, 4), (31, 1), (32, 4)])
def test_every_table_layout_and_staging_class(torch, k, m):
    """k selects the table-group layout (8+2, 16, 16+4, 16+8+4, 16+16 with padding ...), the staging path
    (LDG / 1-D TMA / 2-D tensor map / split phases) and the warp count at compile time: encode, verify and
    reconstruct of ragged stripes against the oracle for one k of every class, with strides on both sides of
    the 512-byte tile (a stride < 512 B takes the fallback without a tensor map)."""
    tot = k + m
    for stride, lens in ((2048 + 16, [2064, 1, 17, 511, 512, 513, 1000, 2063]), (400, [400, 399, 16, 255])):
        n = len(lens)
        lens_a = np.array(lens, dtype=np.uint32)
        data = O.fill_random(n * k * stride, 1000 * k + m + stride)
        P = O.build_matrix(k, m, 0)
        want = O.encode(k, m, P, data, stride, n, lens_a)
        with G.GarageEc(0, k, m) as ec:
            d = dev(torch, data)
            dl = dev(torch, lens_a.astype(np.int32))
            par = torch.full((n * m * stride,), 0xEE, dtype=torch.uint8, device="cuda")
            ec.encode(d, par, stride, n, shard_len=dl)
            got = host(par).reshape(n, m, stride)
            w3 = want.reshape(n, m, stride)
            for s in range(n):
                Lp = (lens[s] + 15) // 16 * 16
                assert np.array_equal(got[s, :, :lens[s]], w3[s, :, :lens[s]]), (k, m, stride, s)
                assert not got[s, :, lens[s]:Lp].any() and (got[s, :, Lp:] == 0xEE).all()
            sh = np.concatenate([data.reshape(n, k, stride), w3], axis=1)
            mm = torch.ones(n, dtype=torch.int32, device="cuda")
            ec.verify(dev(torch, sh.reshape(-1)), mm, stride, n, shard_len=dl)
            assert int(mm.abs().sum()) == 0
            hurt = sh.copy()
            hurt[0, k, 0] ^= 1
            hurt[n - 1, k + m - 1, lens[n - 1] - 1] ^= 0x80
            ec.verify(dev(torch, hurt.reshape(-1)), mm, stride, n, shard_len=dl)
            mmh = host(mm)
            assert mmh[0] == 1 and mmh[n - 1] == 1 << (m - 1) and np.count_nonzero(mmh) == 2
            rng = np.random.default_rng(k * 100 + m)
            present = np.ones((n, tot), dtype=np.uint8)
            for s in range(n):
                present[s, rng.choice(tot, int(rng.integers(1, m + 1)), replace=False)] = 0
            broken = sh.copy()
            broken[present == 0] = 0x77
            bd = dev(torch, broken.reshape(-1))
            st = torch.ones(n, dtype=torch.int32, device="cuda")
            ec.reconstruct(bd, dev(torch, present), stride, n, status=st, shard_len=dl)
            assert int(st.abs().sum()) == 0
            rec = host(bd).reshape(n, tot, stride)
            for s in range(n):
                assert np.array_equal(rec[s, :, :lens[s]], sh[s, :, :lens[s]]), (k, m, stride, s)

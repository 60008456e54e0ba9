"""Full byte vectors anchored outside this repo's oracle (tests/golden/external_anchors.json, generated
by tests/golden/make_external_anchors.py with its own pure-python field arithmetic):
the Backblaze JavaReedSolomon 4+2 worked example (data "ABCD/EFGH/IJKL/MNOP" -> parity
51 52 53 49 / 55 56 57 25, as published) and the Linux-RAID-6 P/Q syndromes of one 1 MiB block.
The CPU oracle (both arms) and the GPU path must reproduce them byte for byte."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import oracle_lib as O  # noqa: E402
import make_external_anchors as MA  # noqa: E402

FIX = json.load(open(os.path.join(HERE, "golden", "external_anchors.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_fixture_is_reproducible_and_matches_the_published_backblaze_numbers():
    assert MA.build() == FIX
    bb = FIX["backblaze_4_2"]
    assert bb["matrix"] == [[27, 28, 18, 20], [28, 27, 20, 18]]
    assert bb["parity_hex"] == ["51525349", "55565725"]  # the figures of the Backblaze write-up
    # the oracle's own matrix construction gives the same rows
    assert O.build_matrix(4, 2, 0).tolist() == bb["matrix"]
    # the data stream of the fixture is the oracle's / the device's stream
    assert sha(O.fill_random(1 << 20, MA.SEED, 0)) == FIX["raid6_1MiB"]["block_sha256"]


def _bb_case():
    bb = FIX["backblaze_4_2"]
    stride = 16
    data = np.zeros((4, stride), dtype=np.uint8)
    for j, sx in enumerate(bb["data"]):
        data[j, :4] = np.frombuffer(sx.encode(), dtype=np.uint8)
    want = np.zeros((2, stride), dtype=np.uint8)
    for i, hx in enumerate(bb["parity_hex"]):
        want[i, :4] = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8)
    return data, want, stride, np.array([4], dtype=np.uint32)


def _raid6_case():
    r = FIX["raid6_1MiB"]
    k, L = r["k"], r["shard_len"]
    stride = (L + 127) // 128 * 128
    blk = O.fill_random(r["block_len"], int(r["seed"], 16), 0)
    data = O.split_block(blk, k, stride)
    return r, k, L, stride, data, np.array(r["matrix"], dtype=np.uint8)


@pytest.mark.parametrize("simd", [False, True])
def test_oracle_reproduces_external_vectors(simd):
    data, want, stride, lens = _bb_case()
    P = O.build_matrix(4, 2, 0)
    got = O.encode(4, 2, P, data.reshape(-1), stride, 1, lens, simd=simd).reshape(2, stride)
    assert np.array_equal(got, want)
    r, k, L, stride, data, M = _raid6_case()
    par = O.encode(k, 2, M, data, stride, 1, np.array([L], dtype=np.uint32), simd=simd).reshape(2, stride)
    assert sha(par[0, :L]) == r["P_sha256"] and sha(par[1, :L]) == r["Q_sha256"]
    assert par[0, :32].tobytes().hex() == r["P_first32"] and par[1, L - 16:L].tobytes().hex() == r["Q_last16"]


@pytest.mark.gpu
def test_gpu_reproduces_external_vectors():
    import torch

    import garage_b200 as G

    data, want, stride, lens = _bb_case()
    with G.GarageEc(0, 4, 2) as ec:
        par = torch.zeros(2 * stride, dtype=torch.uint8, device="cuda")
        ec.encode(torch.from_numpy(data.reshape(-1)).cuda(), par, stride, 1,
                  shard_len=torch.from_numpy(lens.astype(np.int32)).cuda())
        assert np.array_equal(par.cpu().numpy().reshape(2, stride), want)
        # HOST entry point, and the block-level framing ("ABCDEFGHIJKLMNOP" is one 16-byte block)
        hp = np.zeros(2 * stride, dtype=np.uint8)
        ec.encode(data.reshape(-1), hp, stride, 1, shard_len=lens)
        assert np.array_equal(hp.reshape(2, stride), want)
        blk = np.frombuffer(b"ABCDEFGHIJKLMNOP", dtype=np.uint8).copy()
        bp = np.zeros(2 * 128, dtype=np.uint8)
        ec.encode_blocks([blk], bp, 128)
        assert bp.reshape(2, 128)[:, :4].tobytes().hex() == "5152534955565725"
        # lose "ABCD" and "MNOP", rebuild them from the published parity
        sh = np.concatenate([data, want]).copy()
        present = np.array([[0, 1, 1, 0, 1, 1]], dtype=np.uint8)
        sh[[0, 3]] = 0
        d = torch.from_numpy(sh.reshape(-1)).cuda()
        ec.reconstruct(d, torch.from_numpy(present).cuda(), stride, 1,
                       shard_len=torch.from_numpy(lens.astype(np.int32)).cuda())
        got = d.cpu().numpy().reshape(6, stride)
        assert got[0, :4].tobytes() == b"ABCD" and got[3, :4].tobytes() == b"MNOP"
    r, k, L, stride, data, M = _raid6_case()
    with G.GarageEc(0, k, 2, matrix=M) as ec:
        par = torch.zeros(2 * stride, dtype=torch.uint8, device="cuda")
        dl = torch.tensor([L], dtype=torch.int32, device="cuda")
        ec.encode(torch.from_numpy(data).cuda(), par, stride, 1, shard_len=dl)
        got = par.cpu().numpy().reshape(2, stride)
        assert sha(got[0, :L]) == r["P_sha256"] and sha(got[1, :L]) == r["Q_sha256"]
        assert got[0, :32].tobytes().hex() == r["P_first32"] and got[1, :32].tobytes().hex() == r["Q_first32"]
        assert got[0, L - 16:L].tobytes().hex() == r["P_last16"] and got[1, L - 16:L].tobytes().hex() == r["Q_last16"]
        assert not got[:, L:].any()
        # double-disk failure: rebuild data shards 2 and 7 from P and Q, compare with the block
        sh = np.concatenate([data.reshape(k, stride), got]).copy()
        orig = sh.copy()
        sh[[2, 7]] = 0
        present = np.ones((1, k + 2), dtype=np.uint8)
        present[0, [2, 7]] = 0
        d = torch.from_numpy(sh.reshape(-1)).cuda()
        ec.reconstruct(d, torch.from_numpy(present).cuda(), stride, 1, shard_len=dl)
        assert np.array_equal(d.cpu().numpy().reshape(k + 2, stride), orig)

"""An external anchor for the field arithmetic and the encode path: RAID-6.

Linux md RAID-6 (H. P. Anvin, "The mathematics of RAID-6") works in the same field -- GF(2^8) with
polynomial 0x11D and generator {02} -- and defines  P = D_0 ^ D_1 ^ ... ,  Q = sum g^i * D_i.
With the parity matrix [[1,1,..,1],[1,2,4,..,2^(k-1)]] our encode must therefore produce the RAID-6
P and Q syndromes.  The check below computes Q the way the kernel's generic C code does (Horner's
rule, multiply-by-2 as shift + conditional xor 0x1d; no tables), so it shares nothing with the
oracle's log/exp tables or the GPU's product tables."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle_lib as O  # noqa: E402


def raid6_pq(data):
    """data: (k, L) uint8 -> (P, Q) by Horner's rule, highest disk first (lib/raid6/int.uc)"""
    k = data.shape[0]
    p = data[k - 1].copy()
    q = data[k - 1].copy()
    for z in range(k - 2, -1, -1):
        hi = (q & 0x80) != 0
        q = ((q << 1) & 0xFF) ^ np.where(hi, 0x1D, 0).astype(np.uint8)  # multiply by {02}
        q ^= data[z]
        p ^= data[z]
    return p, q


def raid6_matrix(k):
    P = np.ones((2, k), dtype=np.uint8)
    v = 1
    for j in range(k):
        P[1, j] = v
        v = (v << 1) ^ (0x11D if v & 0x80 else 0)
        v &= 0xFF
    return P


@pytest.mark.parametrize("k", [2, 4, 10, 16])
def test_oracle_reproduces_raid6_syndromes(k):
    L = 4099
    data = O.fill_random(k * L, 60 + k).reshape(k, L)
    P = raid6_matrix(k)
    assert list(P[1][:4]) == [1, 2, 4, 8][: min(4, k)]
    par = O.encode(k, 2, P, data.reshape(-1), L, 1).reshape(2, L)
    p, q = raid6_pq(data)
    assert np.array_equal(par[0], p) and np.array_equal(par[1], q)
    par2 = O.encode(k, 2, P, data.reshape(-1), L, 1, simd=True).reshape(2, L)
    assert np.array_equal(par2, par)


@pytest.mark.gpu
@pytest.mark.parametrize("k", [4, 10, 13])
def test_gpu_reproduces_raid6_syndromes_and_recovers_two_disks(k):
    import torch

    import garage_b200 as G

    stride, n = 8192, 5
    data = O.fill_random(n * k * stride, 70 + k).reshape(n, k, stride)
    with G.GarageEc(0, k, 2, matrix=raid6_matrix(k)) as ec:
        par = torch.zeros(n * 2 * stride, dtype=torch.uint8, device="cuda")
        ec.encode(torch.from_numpy(data.reshape(-1)).cuda(), par, stride, n)
        got = par.cpu().numpy().reshape(n, 2, stride)
        for s in range(n):
            p, q = raid6_pq(data[s])
            assert np.array_equal(got[s, 0], p) and np.array_equal(got[s, 1], q)
        # the classic RAID-6 double-disk failure: lose two data disks, rebuild from P and Q
        sh = np.concatenate([data, got], axis=1)
        present = np.ones((n, k + 2), dtype=np.uint8)
        present[:, [0, k - 1]] = 0
        broken = sh.copy()
        broken[:, [0, k - 1]] = 0
        d = torch.from_numpy(broken.reshape(-1)).cuda()
        ec.reconstruct(d, torch.from_numpy(present).cuda(), stride, n)
        assert np.array_equal(d.cpu().numpy().reshape(n, k + 2, stride), sh)

#!/usr/bin/env python
"""Tiny end-to-end pass over every kernel, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck) on the GPU box -- SURVEY.md section 5 'race detection'."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import garage_b200 as G  # noqa: E402
import oracle_lib as O  # noqa: E402

for (k, m) in ((10, 4), (4, 2), (17, 8)):
    tot, stride, n = k + m, 1024 + 16, 6
    lens = np.array([stride, 1, 17, 1000, 640, 33], dtype=np.uint32)
    data = O.fill_random(n * k * stride, k)
    P = O.build_matrix(k, m, 0)
    want = O.encode(k, m, P, data, stride, n, lens)
    with G.GarageEc(0, k, m) as ec:
        d = torch.from_numpy(data).cuda()
        dl = torch.from_numpy(lens.astype(np.int32)).cuda()
        par = torch.zeros(n * m * stride, dtype=torch.uint8, device="cuda")
        ec.encode(d, par, stride, n, shard_len=dl)
        assert np.array_equal(par.cpu().numpy(), want)
        sh = torch.cat([d.view(n, k, stride), par.view(n, m, stride)], dim=1).contiguous()
        orig = sh.clone()
        present = np.ones((n, tot), dtype=np.uint8)
        rng = np.random.default_rng(k)
        for s in range(n):
            present[s, rng.choice(tot, m, replace=False)] = 0
        pd = torch.from_numpy(present).cuda()
        sh[~pd.bool()] = 0
        st = torch.zeros(n, dtype=torch.int32, device="cuda")
        ec.reconstruct(sh.view(-1), pd, stride, n, status=st, shard_len=dl)
        for s in range(n):
            assert torch.equal(sh[s, :, : lens[s]], orig[s, :, : lens[s]])
        mm = torch.ones(n, dtype=torch.int32, device="cuda")
        ec.verify(sh.view(-1), mm, stride, n, shard_len=dl)
        assert int(mm.abs().sum()) == 0
        sums = torch.zeros(n * tot * 32, dtype=torch.uint8, device="cuda")
        ec.shard_sums(sh.view(-1), sums, stride, n, tot, shard_len=dl)
        sh[2, 1, 0] ^= 1
        bad = torch.zeros(n * tot, dtype=torch.uint8, device="cuda")
        ec.scrub_repair(sh.view(-1), sums, bad, stride, n, status=st, shard_len=dl)
        assert int(bad.sum()) == 1
        torch.cuda.synchronize()
print("sanitize_small ok")

"""CPU-only: the oracle (scalar and SIMD arms) and the numpy restatement reproduce the frozen
golden parity digests (tests/golden/parity_digests.json, made by tests/golden/make_golden.py)."""
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
import make_golden as MG  # noqa: E402

sys.path.insert(0, os.path.join(O.ROOT, "oracle"))
import rs_oracle_np as NP  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "parity_digests.json")))["cases"]


@pytest.mark.parametrize("name", sorted(GOLD))
def test_oracle_matches_golden(name):
    g = GOLD[name]
    k, m, kind, n = g["k"], g["m"], g["kind"], g["n"]
    bl = g["block_lens"][0] if len(set(g["block_lens"])) == 1 else None
    data, L, stride, lens = MG.make_case(k, m, kind, n, bl)
    assert lens == g["block_lens"] and stride == g["stride"]
    assert hashlib.sha256(data.tobytes()).hexdigest() == g["data_sha256"]
    P = O.build_matrix(k, m, kind)
    for simd in (False, True):
        par = O.encode(k, m, P, data, stride, n, L, simd=simd)
        assert hashlib.sha256(par.tobytes()).hexdigest() == g["parity_sha256"]
        assert par[:64].tobytes().hex() == g["parity_first64"]
    # numpy restatement on the first stripe (it is slow)
    d0 = data[: k * stride].reshape(k, stride)[:, : L[0]]
    p0 = par[: m * stride].reshape(m, stride)[:, : L[0]]
    assert np.array_equal(NP.encode(NP.build_matrix(k, m, kind), d0), p0)

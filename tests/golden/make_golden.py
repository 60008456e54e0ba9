"""Generates tests/golden/parity_digests.json from the scalar CPU oracle (oracle/rs_oracle.c).

The reference has no RS implementation, so these are digests of the
oracle's own output, frozen so that any later change to oracle OR kernels is caught.
Inputs: block s = bytes [s*B, (s+1)*B) of the splitmix64 stream with seed
0x6761726167650010; shard layout, stride = shard_len rounded up to
128 B; parity buffer zero outside [0, shard_len).

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402

SEED = 0x6761726167650010
CASES = [  # (name, k, m, kind, n_blocks, block_len)
    ("config0_rs4_2_64x1MiB", 4, 2, 0, 64, 1 << 20),
    ("rs10_4_8x1MiB", 10, 4, 0, 8, 1 << 20),
    ("rs6_3_8x1MiB", 6, 3, 0, 8, 1 << 20),
    ("rs10_4_cauchy_4x1MiB", 10, 4, 1, 4, 1 << 20),
    ("rs10_4_short_blocks", 10, 4, 0, 6, None),
]
SHORT = [3073, 65536, 1, 1048575, 500000, 12345]


def make_case(k, m, kind, n, block_len):
    lens = [block_len] * n if block_len else SHORT[:n]
    L = [O.lib().rs_oracle_shard_len(b, k) for b in lens]
    stride = (max(L) + 127) // 128 * 128
    data = np.zeros(n * k * stride, dtype=np.uint8)
    off = 0
    for s, b in enumerate(lens):
        blk = O.fill_random(b, SEED, off)
        off += (b + 7) // 8 * 8
        data[s * k * stride:(s + 1) * k * stride] = O.split_block(blk, k, stride)
    return data, np.array(L, dtype=np.uint32), stride, lens


def main():
    out = {"_comment": __doc__.strip().splitlines()[0], "seed": hex(SEED), "cases": {}}
    for name, k, m, kind, n, bl in CASES:
        data, L, stride, lens = make_case(k, m, kind, n, bl)
        P = O.build_matrix(k, m, kind)
        par = O.encode(k, m, P, data, stride, n, L)
        out["cases"][name] = {
            "k": k, "m": m, "kind": kind, "n": n, "block_lens": lens, "stride": stride,
            "data_sha256": hashlib.sha256(data.tobytes()).hexdigest(),
            "parity_sha256": hashlib.sha256(par.tobytes()).hexdigest(),
            "parity_first64": par[:64].tobytes().hex(),
        }
        print(name, out["cases"][name]["parity_sha256"])
    json.dump(out, open(os.path.join(HERE, "parity_digests.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

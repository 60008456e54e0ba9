"""Generates tests/golden/external_anchors.json: full byte vectors anchored OUTSIDE this repo's
oracle (the reference itself holds no RS vector: SURVEY.md section 0.1, doc/book/design/goals.md:27).

This script shares no code with oracle/ or garage_b200/: GF(2^8)/0x11D products are carry-less
shift-and-xor, the Vandermonde-systematic matrix is built and inverted right here, and the
synthetic block is the splitmix64 stream written out in numpy.

1. backblaze_4_2   The 4+2 example of Backblaze's JavaReedSolomon write-up: data shards "ABCD",
                   "EFGH", "IJKL", "MNOP"; the published parity rows of the Vandermonde-systematic
                   4+2 code are (decimal) 27 28 18 20 / 28 27 20 18 -- the script ASSERTS that its
                   own construction reproduces exactly those rows, then applies them.
2. raid6_1MiB      One RS(10,2) stripe of a 1 MiB block with the Linux md RAID-6 syndromes
                   (H. P. Anvin, "The mathematics of RAID-6": same field, generator {02}):
                   P = xor of the data shards, Q = sum 2^j * D_j by Horner's rule.  With the
                   parity matrix [[1,..,1],[1,2,4,..]] encode must produce exactly P and Q.

    python tests/golden/make_external_anchors.py        (rewrites the JSON; the tests check that a
                                                         fresh run reproduces the committed file)
"""
import hashlib
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 0x6761726167650010
B = 1 << 20


def gmul(a, b):
    """carry-less multiply modulo x^8+x^4+x^3+x^2+1"""
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        if a & 0x100:
            a ^= 0x11D
        b >>= 1
    return r


def ginv(a):
    r = 1
    for _ in range(254):  # a^254 = a^-1
        r = gmul(r, a)
    return r


def gpow(a, e):
    r = 1
    for _ in range(e):
        r = gmul(r, a)
    return r


def vandermonde_systematic(k, m):
    V = [[gpow(r, c) for c in range(k)] for r in range(k + m)]  # 0^0 = 1
    A = [row[:] + [1 if i == j else 0 for j in range(k)] for i, row in enumerate(V[:k])]
    for c in range(k):  # Gauss-Jordan on [V_top | I]
        piv = next(r for r in range(c, k) if A[r][c])
        A[c], A[piv] = A[piv], A[c]
        iv = ginv(A[c][c])
        A[c] = [gmul(x, iv) for x in A[c]]
        for r in range(k):
            if r != c and A[r][c]:
                f = A[r][c]
                A[r] = [x ^ gmul(f, y) for x, y in zip(A[r], A[c])]
    inv = [row[k:] for row in A]
    P = []
    for i in range(m):
        P.append([0] * k)
        for j in range(k):
            acc = 0
            for x in range(k):
                acc ^= gmul(V[k + i][x], inv[x][j])
            P[i][j] = acc
    return P


def splitmix_bytes(nbytes, seed, byte_off=0):
    """the synthetic-data stream of SURVEY.md 8(d): word i = splitmix64 finaliser of seed + (i+1)*golden"""
    first = byte_off // 8
    n = (nbytes + 7) // 8
    i = np.arange(first + 1, first + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)[:nbytes].copy()


def xtime(v):
    hi = (v & 0x80) != 0
    return ((v << 1) & 0xFF).astype(np.uint8) ^ np.where(hi, 0x1D, 0).astype(np.uint8)


def build():
    out = {"_comment": __doc__.strip().splitlines()[0]}
    # ---- 1. Backblaze 4+2
    P = vandermonde_systematic(4, 2)
    assert P == [[27, 28, 18, 20], [28, 27, 20, 18]], P  # the published rows
    data = [b"ABCD", b"EFGH", b"IJKL", b"MNOP"]
    par = []
    for i in range(2):
        row = bytearray(4)
        for t in range(4):
            acc = 0
            for j in range(4):
                acc ^= gmul(P[i][j], data[j][t])
            row[t] = acc
        par.append(bytes(row))
    out["backblaze_4_2"] = {"k": 4, "m": 2, "matrix": P, "data": [d.decode() for d in data],
                            "parity_hex": [p.hex() for p in par]}
    # ---- 2. RAID-6 P/Q of one 1 MiB block, RS(10,2)
    k = 10
    L = (B + k - 1) // k
    blk = splitmix_bytes(B, SEED, 0)
    sh = np.zeros((k, L), dtype=np.uint8)
    flat = np.zeros(k * L, dtype=np.uint8)
    flat[:B] = blk
    sh[:] = flat.reshape(k, L)
    p = sh[k - 1].copy()
    q = sh[k - 1].copy()
    for z in range(k - 2, -1, -1):
        q = xtime(q) ^ sh[z]
        p ^= sh[z]
    mat = [[1] * k, []]
    v = 1
    for j in range(k):
        mat[1].append(v)
        v = gmul(v, 2)
    out["raid6_1MiB"] = {"k": k, "m": 2, "matrix": mat, "seed": hex(SEED), "block_len": B, "shard_len": L,
                         "block_sha256": hashlib.sha256(blk.tobytes()).hexdigest(),
                         "P_sha256": hashlib.sha256(p.tobytes()).hexdigest(),
                         "Q_sha256": hashlib.sha256(q.tobytes()).hexdigest(),
                         "P_first32": p[:32].tobytes().hex(), "Q_first32": q[:32].tobytes().hex(),
                         "P_last16": p[-16:].tobytes().hex(), "Q_last16": q[-16:].tobytes().hex()}
    return out


if __name__ == "__main__":
    d = build()
    json.dump(d, open(os.path.join(HERE, "external_anchors.json"), "w"), indent=1)
    print(json.dumps(d["backblaze_4_2"]))
    print(d["raid6_1MiB"]["P_sha256"], d["raid6_1MiB"]["Q_sha256"])

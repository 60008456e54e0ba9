"""Independent numpy restatement of the RS block-path arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/rs_oracle.h).  PARITY UNPINNED by the reference:
deuxfleurs-org/garage has no Reed-Solomon code (doc/book/design/goals.md:27, SURVEY.md
section 0.1); this second implementation exists so that the C oracle is cross-checked
by something that shares none of its code paths:

  * multiplication is shift-and-xor ("Russian peasant") reduction by 0x11D, never log/exp;
  * the Vandermonde-systematic matrix is obtained by solving  X * V[0:k] = V[k:k+m]
    with Gaussian elimination on the transposed system, not by inverting then multiplying;
  * decode builds ONE composed (missing x k) matrix from the k survivors instead of the
    oracle's two steps (recover data, then re-encode parity).

Framing follows src/block/block.rs:85-96 (which bytes are encoded) and
src/api/s3/put.rs:583-617 (block sizes; short last block).
"""
from __future__ import annotations

import numpy as np

POLY = 0x11D
VANDERMONDE = 0
CAUCHY = 1


def gf_mul_scalar(a: int, b: int) -> int:
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        if a & 0x100:
            a ^= POLY
        b >>= 1
    return r


def _build_mul_table() -> np.ndarray:
    t = np.zeros((256, 256), dtype=np.uint8)
    for a in range(256):
        for b in range(a, 256):
            v = gf_mul_scalar(a, b)
            t[a, b] = v
            t[b, a] = v
    return t


MUL = _build_mul_table()


def gf_inv_scalar(a: int) -> int:
    assert a != 0
    # a^254 by square-and-multiply
    r, base, e = 1, a, 254
    while e:
        if e & 1:
            r = int(MUL[r, base])
        base = int(MUL[base, base])
        e >>= 1
    return r


def gf_pow(a: int, n: int) -> int:
    r = 1
    for _ in range(n):
        r = int(MUL[r, a])
    return r


def gf_matmul(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    A = np.asarray(A, dtype=np.uint8)
    B = np.asarray(B, dtype=np.uint8)
    out = np.zeros((A.shape[0], B.shape[1]), dtype=np.uint8)
    for x in range(A.shape[1]):
        out ^= MUL[A[:, x][:, None], B[x, :][None, :]]
    return out


def gf_solve(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """Solve A X = B over GF(2^8) (A square, non-singular)."""
    A = np.array(A, dtype=np.uint8)
    B = np.array(B, dtype=np.uint8)
    n = A.shape[0]
    for c in range(n):
        piv = next((r for r in range(c, n) if A[r, c]), None)
        if piv is None:
            raise np.linalg.LinAlgError("singular over GF(2^8)")
        if piv != c:
            A[[c, piv]] = A[[piv, c]]
            B[[c, piv]] = B[[piv, c]]
        inv = gf_inv_scalar(int(A[c, c]))
        A[c] = MUL[inv, A[c]]
        B[c] = MUL[inv, B[c]]
        for r in range(n):
            if r != c and A[r, c]:
                f = int(A[r, c])
                A[r] ^= MUL[f, A[c]]
                B[r] ^= MUL[f, B[c]]
    return B


def build_matrix(k: int, m: int, kind: int = VANDERMONDE) -> np.ndarray:
    if kind == CAUCHY:
        return np.array(
            [[gf_inv_scalar((k + i) ^ j) for j in range(k)] for i in range(m)], dtype=np.uint8
        )
    V = np.array([[gf_pow(r, c) for c in range(k)] for r in range(k + m)], dtype=np.uint8)
    top, bot = V[:k], V[k:]
    # P * top = bot  <=>  top^T * P^T = bot^T
    return gf_solve(top.T, bot.T).T.copy()


def generator(k: int, m: int, P: np.ndarray) -> np.ndarray:
    return np.concatenate([np.eye(k, dtype=np.uint8), np.asarray(P, dtype=np.uint8)], axis=0)


def shard_len(block_len: int, k: int) -> int:
    return (block_len + k - 1) // k


def split_block(block: bytes | np.ndarray, k: int) -> np.ndarray:
    b = np.frombuffer(bytes(block), dtype=np.uint8) if not isinstance(block, np.ndarray) else block
    L = shard_len(len(b), k)
    out = np.zeros((k, L), dtype=np.uint8)
    flat = out.reshape(-1)
    flat[: len(b)] = b
    return out


def encode(P: np.ndarray, data: np.ndarray) -> np.ndarray:
    """data: (k, L) uint8 -> parity (m, L)."""
    P = np.asarray(P, dtype=np.uint8)
    out = np.zeros((P.shape[0], data.shape[1]), dtype=np.uint8)
    for i in range(P.shape[0]):
        for j in range(P.shape[1]):
            out[i] ^= MUL[int(P[i, j])][data[j]]
    return out


def decode_matrix(k: int, m: int, P: np.ndarray, present: np.ndarray):
    """Composed matrix D (missing x k) s.t. missing_shards = D * survivors, survivors being
    the first k present shard indices.  Returns (D, survivors, missing) or None if < k."""
    present = np.asarray(present).astype(bool)
    surv = [i for i in range(k + m) if present[i]][:k]
    if len(surv) < k:
        return None
    missing = [i for i in range(k + m) if not present[i]]
    G = generator(k, m, P)
    S = G[surv]  # k x k : survivors = S * data
    # rows_missing = G[missing] * S^-1  <=>  D * S = G[missing]  <=> S^T D^T = G[missing]^T
    if not missing:
        return np.zeros((0, k), dtype=np.uint8), surv, missing
    D = gf_solve(S.T, G[missing].T).T.copy()
    return D, surv, missing


def reconstruct(k: int, m: int, P: np.ndarray, shards: np.ndarray, present: np.ndarray) -> bool:
    """shards: (k+m, L) modified in place. Returns False if unrecoverable."""
    r = decode_matrix(k, m, P, present)
    if r is None:
        return False
    D, surv, missing = r
    if missing:
        shards[missing] = encode(D, shards[surv])
    return True


def verify(k: int, m: int, P: np.ndarray, shards: np.ndarray) -> int:
    par = encode(P, shards[:k])
    mm = 0
    for i in range(m):
        if not np.array_equal(par[i], shards[k + i]):
            mm |= 1 << i
    return mm


_M64 = (1 << 64) - 1


def fill_random(n: int, seed: int, offset: int = 0) -> np.ndarray:
    """Same splitmix64 counter stream as rs_oracle_fill_random (offset multiple of 8)."""
    idx0 = offset // 8
    nw = (n + 7) // 8
    with np.errstate(over="ignore"):
        i = np.arange(idx0 + 1, idx0 + nw + 1, dtype=np.uint64)
        z = np.uint64(seed & _M64) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.astype("<u8").view(np.uint8)[:n].copy()

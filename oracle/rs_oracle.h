/*
 * rs_oracle.h -- CPU ORACLE for the Garage erasure-coding block path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg may link, import or execute it, and there only as the checker.
 *
 * PARITY UNPINNED.  The reference (deuxfleurs-org/garage v1.2.0 @3f4ab3a4) contains
 * no Reed-Solomon code, no RS crate in Cargo.lock and no golden vectors for parity
 * (erasure coding is a documented non-goal, doc/book/design/goals.md:27; SURVEY.md
 * section 0.1).  This file therefore *is* the normative definition; it restates the
 * published construction shared by the reed-solomon-erasure crate / klauspost
 * reedsolomon / Backblaze JavaReedSolomon:
 *
 *   field    GF(2^8), reduction polynomial 0x11D, generator alpha = 2
 *   kind 0   Vandermonde-systematic: V[r][c] = r^c, G = V * inv(V[0..k)), P = G[k..k+m)
 *   kind 1   Cauchy: P[i][j] = 1 / ((k+i) xor j)
 *   framing  one stripe = one (post-compression) block, src/block/block.rs:85-96;
 *            shard_len = ceil(len/k), data shard j = bytes [j*shard_len,(j+1)*shard_len)
 *            zero-padded at the tail (block sizes: src/api/s3/put.rs:583-617)
 *   encode   parity[i][t] = XOR_j P[i][j] * data[j][t]
 *   decode   first k present rows of [I;P] -> invert (Gauss-Jordan) -> missing data
 *            shards; missing parity shards re-encoded from the completed data.
 *
 * It is pinned by (a) the survey-time known-answer vectors of SURVEY.md section 8(c)
 * (tests/golden/kat.json), (b) an independent numpy restatement
 * (oracle/rs_oracle_np.py) that must agree byte for byte, (c) MDS and round-trip
 * property tests.
 */
#ifndef RS_ORACLE_H
#define RS_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RS_ORACLE_VANDERMONDE 0
#define RS_ORACLE_CAUCHY 1

/* field */
uint8_t rs_oracle_gf_mul(uint8_t a, uint8_t b);
uint8_t rs_oracle_gf_inv(uint8_t a); /* a != 0 */
uint8_t rs_oracle_gf_exp(int i);     /* alpha^i, i in [0,510) */
uint8_t rs_oracle_gf_log(uint8_t a); /* a != 0 */

/* parity rows P (m x k, row-major).  returns 0, or -1 on bad (k,m,kind). */
int rs_oracle_build_matrix(int k, int m, int kind, uint8_t *P);

/* in-place inversion of an n x n matrix (row-major). returns 0, -1 if singular */
int rs_oracle_invert(uint8_t *M, int n);

/* batch geometry shared with include/garage_ec.h:
 *   data   : n * k * stride bytes, shard j of stripe s at (s*k + j) * stride
 *   parity : n * m * stride bytes, row  i of stripe s at (s*m + i) * stride
 *   shard_len[s] <= stride valid bytes per shard (NULL -> stride for all)      */
void rs_oracle_encode(int k, int m, const uint8_t *P, const uint8_t *data, uint8_t *parity,
                      const uint32_t *shard_len, size_t stride, size_t n);

/* shards: n * (k+m) * stride, shard i of stripe s at (s*(k+m)+i)*stride; present[s*(k+m)+i]
 * in {0,1}.  Rewrites every absent shard in place.  status[s] = 0 ok / -1 unrecoverable
 * (< k present; shards left untouched).  returns number of unrecoverable stripes.        */
size_t rs_oracle_reconstruct(int k, int m, const uint8_t *P, uint8_t *shards,
                             const uint8_t *present, int32_t *status,
                             const uint32_t *shard_len, size_t stride, size_t n);

/* scrub: mismatch[s] bit i set iff recomputed parity row i differs from stored one. */
void rs_oracle_verify(int k, int m, const uint8_t *P, const uint8_t *shards, uint32_t *mismatch,
                      const uint32_t *shard_len, size_t stride, size_t n);

/* framing helpers (block.rs:85-96 decides which bytes; put.rs:611-615 short last block) */
uint32_t rs_oracle_shard_len(uint32_t block_len, int k);
/* split one block into k zero-padded shards at dst + j*stride (only shard_len bytes written) */
void rs_oracle_split_block(const uint8_t *block, uint32_t block_len, int k, uint8_t *dst,
                           size_t stride);
/* inverse: concatenate the k data shards back into block_len bytes */
void rs_oracle_join_block(const uint8_t *shards, uint32_t block_len, int k, size_t stride,
                          uint8_t *block);

/* synthetic input generator shared by tests and bench:
 * splitmix64 keyed by (seed, 8-byte word index); fills dst[0..len) as the bytes at absolute
 * offset `offset` (multiple of 8) of the infinite stream. */
void rs_oracle_fill_random(uint8_t *dst, size_t len, uint64_t seed, uint64_t offset);

#ifdef __cplusplus
}
#endif
#endif

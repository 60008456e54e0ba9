// blake2b.h -- BLAKE2b (RFC 7693) with a 64-byte digest, truncated to the first 32 bytes:
// exactly what Garage calls `blake2sum` and uses as the content address of a block
// (src/util/data.rs:130-138: Blake2b512::new(); update; finalize()[..32]) and to verify a
// plain block on read / scrub (src/block/block.rs:69-83).  One implementation for host and
// device (`GEC_HD`), used by the per-shard integrity of the EC path:
// a storage node holds 1/k of a block, so the whole-block hash cannot be checked locally;
// every shard therefore carries its own blake2sum.
//
// Pinned: RFC 7693 test vector + python hashlib.blake2b(...).digest()[:32] (tests/test_blake2.py).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define GEC_HD __host__ __device__ __forceinline__
#else
#define GEC_HD inline
#endif

namespace garage_ec {

struct Blake2bState {
    uint64_t h[8];
    uint64_t t;  // bytes compressed so far (messages < 2^64 bytes)
};

GEC_HD uint64_t b2_rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

#define GEC_B2_G(a, b, c, d, x, y) \
    do {                           \
        a = a + b + (x);           \
        d = b2_rotr(d ^ a, 32);    \
        c = c + d;                 \
        b = b2_rotr(b ^ c, 24);    \
        a = a + b + (y);           \
        d = b2_rotr(d ^ a, 16);    \
        c = c + d;                 \
        b = b2_rotr(b ^ c, 63);    \
    } while (0)

#define GEC_B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                    \
        GEC_B2_G(v[0], v[4], v[8], v[12], m[s0], m[s1]);                                    \
        GEC_B2_G(v[1], v[5], v[9], v[13], m[s2], m[s3]);                                    \
        GEC_B2_G(v[2], v[6], v[10], v[14], m[s4], m[s5]);                                   \
        GEC_B2_G(v[3], v[7], v[11], v[15], m[s6], m[s7]);                                   \
        GEC_B2_G(v[0], v[5], v[10], v[15], m[s8], m[s9]);                                   \
        GEC_B2_G(v[1], v[6], v[11], v[12], m[s10], m[s11]);                                 \
        GEC_B2_G(v[2], v[7], v[8], v[13], m[s12], m[s13]);                                  \
        GEC_B2_G(v[3], v[4], v[9], v[14], m[s14], m[s15]);                                  \
    } while (0)

// one compression of a 128-byte block given as 16 little-endian words
GEC_HD void blake2b_compress(Blake2bState &S, const uint64_t (&m)[16], bool last)
{
    const uint64_t iv0 = 0x6a09e667f3bcc908ull, iv1 = 0xbb67ae8584caa73bull, iv2 = 0x3c6ef372fe94f82bull,
                   iv3 = 0xa54ff53a5f1d36f1ull, iv4 = 0x510e527fade682d1ull, iv5 = 0x9b05688c2b3e6c1full,
                   iv6 = 0x1f83d9abfb41bd6bull, iv7 = 0x5be0cd19137e2179ull;
    uint64_t v[16];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) v[i] = S.h[i];
    v[8] = iv0;
    v[9] = iv1;
    v[10] = iv2;
    v[11] = iv3;
    v[12] = iv4 ^ S.t;
    v[13] = iv5;  // high word of the counter is always 0 here
    v[14] = last ? ~iv6 : iv6;
    v[15] = iv7;
    GEC_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    GEC_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    GEC_B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    GEC_B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    GEC_B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    GEC_B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    GEC_B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    GEC_B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    GEC_B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    GEC_B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    GEC_B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    GEC_B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) S.h[i] ^= v[i] ^ v[i + 8];
}

GEC_HD void blake2b_init512(Blake2bState &S)
{
    S.h[0] = 0x6a09e667f3bcc908ull ^ 0x01010040ull;  // digest 64 bytes, no key, fanout 1, depth 1
    S.h[1] = 0xbb67ae8584caa73bull;
    S.h[2] = 0x3c6ef372fe94f82bull;
    S.h[3] = 0xa54ff53a5f1d36f1ull;
    S.h[4] = 0x510e527fade682d1ull;
    S.h[5] = 0x9b05688c2b3e6c1full;
    S.h[6] = 0x1f83d9abfb41bd6bull;
    S.h[7] = 0x5be0cd19137e2179ull;
    S.t = 0;
}

// Host: Garage's blake2sum -- BLAKE2b-512 of data[0..len), first 32 bytes to out.
inline void blake2sum_host(const uint8_t *data, size_t len, uint8_t out[32])
{
    Blake2bState S;
    blake2b_init512(S);
    uint64_t m[16];
    size_t off = 0;
    while (len - off > 128) {
        memcpy(m, data + off, 128);  // little-endian host
        S.t += 128;
        blake2b_compress(S, m, false);
        off += 128;
    }
    uint8_t last[128];
    memset(last, 0, sizeof(last));
    if (len > off) memcpy(last, data + off, len - off);
    memcpy(m, last, 128);
    S.t += len - off;
    blake2b_compress(S, m, true);
    memcpy(out, S.h, 32);
}

// adler8 shard tag, host side (same definition as adler8_shards_kernel in rs_kernels.cuh)
inline void adler8_host(const uint8_t *data, size_t len, uint8_t out[32])
{
    const size_t seg = (((len + 7) / 8) + 15) / 16 * 16;
    for (int s = 0; s < 8; s++) {
        const size_t start = (size_t)s * seg;
        uint32_t a = 1, b = 0;
        if (start < len) {
            const size_t end = len < start + seg ? len : start + seg;
            size_t i = start;
            while (i < end) {  // 5552: largest run whose sums cannot overflow 32 bits (zlib's NMAX)
                const size_t stop = end - i > 5552 ? i + 5552 : end;
                for (; i < stop; i++) {
                    a += data[i];
                    b += a;
                }
                a %= 65521;
                b %= 65521;
            }
        }
        const uint32_t v = (b << 16) | a;
        out[4 * s + 0] = (uint8_t)v;
        out[4 * s + 1] = (uint8_t)(v >> 8);
        out[4 * s + 2] = (uint8_t)(v >> 16);
        out[4 * s + 3] = (uint8_t)(v >> 24);
    }
}

}  // namespace garage_ec

// rs_kernels.cuh -- sm_100a kernels of the Garage erasure-coding block path.
//
// What they compute (normative definition: DESIGN.md "Arithmetic"):
//     out[i][t] = XOR_j C[i][j] * src[j][t]         over GF(2^8)/0x11D, byte-wise
// for rows i of a small coefficient matrix C (the parity rows P for encode/verify, a
// per-stripe composed decode matrix for reconstruct) and the k source shards of a stripe.
// Reference call sites this replaces work at: BlockManager::rpc_put_block
// (src/block/manager.rs:366-408), rpc_get_raw_block_internal (manager.rs:276-339),
// BlockResyncManager::resync_block (src/block/resync.rs:460-500), DataBlock::verify /
// ScrubWorker::work (src/block/block.rs:69-83, src/block/repair.rs:438-490).
//
// Design (DESIGN.md "Kernels"):
//  * HBM-bound byte work, no tensor cores.  The per-byte GF multiply-accumulate for up to 4
//    output rows is ONE shared-memory lookup: T_j[x] = {C0j*x, C1j*x, C2j*x, C3j*x} packed in
//    a 32-bit word, so a data byte costs one LDS + one XOR for all four rows.
//  * Conflict-free lookups at 16 KB per table ("sub-warp interleaving").  Random byte indices
//    into a plain 256-word table serialise ~3.5-way, lane-private replicas (32 KB/table) do
//    not fit k = 10.  Instead G = 32/R tables share one 32 KB "group": table j = G*t + sub is
//    replicated R times in banks [sub*R, sub*R + R) (word t*8192 + x*32 + sub*R + g).  The
//    warp is cut in G sub-warps of R lanes; at any lookup instruction sub-warp q works on
//    source G*t + (phase xor q), i.e. every sub-warp is in a different table of the group and
//    every lane owns its bank: word = t*8192 + x*32 + (lane xor phase*R).  The source order is
//    simply permuted per sub-warp when the data is LOADED (slot u holds source u xor q), which
//    costs nothing because XOR-accumulation commutes.  RS(10,4): G = 2, 5 groups = 160 KB,
//    1.0 wavefronts per LDS (the first version, plain R = 16 replication, measured 2.07).
//  * log/antilog tables sit in __constant__ memory and are copied to shared memory; they are
//    only used to BUILD the product tables (once per launch for encode/verify, once per
//    change of erasure pattern for reconstruct), never in the streaming loop.
//  * Streaming loop: one 16-byte column of all k shards per thread (coalesced 512 B per warp
//    per shard, ld.global.nc.L1::no_allocate.v4), 16 lookups per shard, 4x4 byte transposes
//    with PRMT, one 16-byte st.global.cs per output row; the next column's vectors are
//    prefetched into registers while the current one is processed.
//  * Persistent grid: one CTA per SM (tables fill shared memory), work items handed to warps
//    (encode/verify) or to CTAs through an atomic counter (reconstruct).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "blake2b.h"
#include "gf256.h"

namespace garage_ec {

// Launch shapes per mode (1 CTA per SM; tables fill shared memory).  PIPE = prefetch the next
// column's k vectors into registers while the current one is processed.  Tuned on B200
// (profiles/r01_variants.md): encode/reconstruct like fewer, fatter threads with the prefetch;
// verify (k+m loads per column, no stores) likes more threads.
#ifndef GEC_NT_ENC
#define GEC_NT_ENC 512
#endif
#ifndef GEC_PIPE_ENC
#define GEC_PIPE_ENC 1
#endif
#ifndef GEC_NT_PLAN
#define GEC_NT_PLAN 512
#endif
#ifndef GEC_PIPE_PLAN
#define GEC_PIPE_PLAN 1
#endif
#ifndef GEC_NT_VER
#define GEC_NT_VER 1024
#endif
#ifndef GEC_PIPE_VER
#define GEC_PIPE_VER 0
#endif
constexpr int kMaxK = 32;
constexpr int kMaxM = 8;
constexpr int kRowsPerPass = 4;  // output rows packed in one 32-bit table word
constexpr uint32_t kGroupBytes = 32768;  // one table group: 256 rows x 32 banks x 4 B

// tables per 32 KB group for k sources: smallest G (least padding) whose groups need <= 192 KB,
// so that >= 32 KB of the SM's 228 KB stay L1 (with 224 KB of tables the global loads starve:
// k = 7 measured 0.65 of peak against 0.92 for k = 6)
#ifndef GEC_MIN_BLOCKS
#define GEC_MIN_BLOCKS 1  // CTAs per SM the streaming kernel is sized for (tuning experiments only)
#endif
__host__ __device__ constexpr int log2_group_for_k(int k)
{
#ifdef GEC_FORCE_LG
    return GEC_FORCE_LG;
#endif
    return k <= 6 ? 0 : (k <= 12 ? 1 : (k <= 24 ? 2 : 3));
}
__host__ __device__ constexpr int slots_for_k(int k)
{
    return ((k + (1 << log2_group_for_k(k)) - 1) >> log2_group_for_k(k)) << log2_group_for_k(k);
}

// GF(2^8) antilog / log tables, pinned in constant memory.  Kernels that need general products
// (decode planning) stage them into shared memory with warp-UNIFORM 16-byte constant loads:
// a per-lane byte index into __constant__ memory serialises 32-way in the constant cache (the
// first version of rs_plan_kernel spent 30 of its 40 us there).
struct GfTables {
    uint8_t exp[512];
    uint8_t log[256];
};
__constant__ __align__(16) GfTables c_gf = {GARAGE_EC_GF_EXP_INIT, GARAGE_EC_GF_LOG_INIT};

__device__ __forceinline__ void stage_gf_tables(uint8_t *s_gf /* 768 B, 16-byte aligned */)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(&c_gf);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t i = warp; i < sizeof(GfTables) / 16; i += nwarps) {
        const uint4 v = src[i];  // same address in every lane: one constant-cache broadcast
        if (lane < 4)
            reinterpret_cast<uint32_t *>(s_gf)[i * 4 + lane] = lane == 0 ? v.x : (lane == 1 ? v.y : (lane == 2 ? v.z : v.w));
    }
}

enum ApplyMode { kModeEncode = 0, kModePlan = 1, kModeVerify = 2 };

// Per-stripe decode plan, produced by rs_plan_kernel, consumed by rs_apply_kernel<kModePlan>.
struct __align__(16) StripePlan {
    unsigned long long key_present;  // bit i: shard i present
    unsigned long long key_out;      // bit i: shard i is rebuilt
    uint8_t nrows;                   // number of shards rebuilt (0..m)
    uint8_t unrecoverable;           // 1 if < k present
    uint8_t pad[6];
    uint8_t surv[kMaxK];           // source shard indices (first k present)
    uint8_t out_idx[kMaxM];        // rebuilt shard indices
    uint8_t coef[kMaxM][kMaxK];  // out[r] = XOR_j coef[r][j] * shard[surv[j]]
};

struct ApplyParams {
    const uint8_t *src;            // stripe s source base = src + s*src_pitch
    uint8_t *dst;                  // stripe s output base = dst + s*dst_pitch
    unsigned long long src_pitch;  // bytes
    unsigned long long dst_pitch;
    const uint32_t *shard_len;  // device, nullable (=> stride)
    const StripePlan *plan;     // kModePlan
    uint32_t *mismatch;         // kModeVerify
    uint32_t *counter;          // kModePlan dynamic scheduler
    uint32_t stride;            // bytes between shards
    uint32_t n;                 // stripes
    uint32_t k;                 // sources
    uint32_t rows;              // outputs this pass (<= 4), uniform modes
    uint32_t row_off;           // first output row of this pass
    uint32_t items_per_stripe;  // ceil(ceil(stride/16)/32), uniform modes
    uint32_t row_bytes;         // 128: bytes per table row (a register operand keeps x*128 an IMAD)
    uint8_t coef[kRowsPerPass * kMaxK];  // uniform modes: coef[i*k + j]
};

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ uint4 ldg_stream(const void *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void *p, const uint4 &v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
// zero the bytes at positions >= nbytes (0 < nbytes < 16) of a 16-byte vector
__device__ __forceinline__ uint4 mask_tail(uint4 v, uint32_t nbytes)
{
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int lo = 4 * i;
        if ((int)nbytes <= lo) w[i] = 0;
        else if ((int)nbytes < lo + 4) w[i] &= 0xffffffffu >> (8 * (lo + 4 - (int)nbytes));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 4x4 byte transpose: a[p] holds {row0,row1,row2,row3} bytes of byte column p;
// returns r[i] = bytes of row i for columns 0..3.
__device__ __forceinline__ void transpose4x4(const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                             const uint32_t a3, uint32_t &r0, uint32_t &r1,
                                             uint32_t &r2, uint32_t &r3)
{
    const uint32_t t0 = __byte_perm(a0, a1, 0x5140);  // a0.b0 a1.b0 a0.b1 a1.b1
    const uint32_t t1 = __byte_perm(a2, a3, 0x5140);
    const uint32_t t2 = __byte_perm(a0, a1, 0x7362);  // a0.b2 a1.b2 a0.b3 a1.b3
    const uint32_t t3 = __byte_perm(a2, a3, 0x7362);
    r0 = __byte_perm(t0, t1, 0x5410);
    r1 = __byte_perm(t0, t1, 0x7632);
    r2 = __byte_perm(t2, t3, 0x5410);
    r3 = __byte_perm(t2, t3, 0x7632);
}

// ------------------------------------------------------------------ shared memory carve-up
// Decode plan of one stripe as the streaming kernel needs it (two slots: current / next).
struct PlanSlot {
    unsigned long long key_present, key_out;
    uint32_t sid;   // stripe index (>= n: no more work)
    int32_t rows;   // rows to produce in this pass (<= 0: nothing to do)
    uint32_t chunk_next;  // next unclaimed 32-column chunk of this stripe (warps claim them dynamically)
    uint32_t pad;
    uint32_t src_off[kMaxK];           // byte offset of source j inside the stripe
    uint32_t dst_off[kRowsPerPass];    // byte offset of output row i inside the stripe
    uint8_t coef[kRowsPerPass * kMaxK];  // coef[i*kMaxK + j]
};
struct SmemLayout {
    // dynamic smem: [table groups: ceil(k/G) * 32 KB][PlanSlot x 2]
    uint32_t *tab;
    PlanSlot *slot;
};
constexpr size_t kSmemAux = 2 * sizeof(PlanSlot) + 16;
__host__ __device__ inline size_t smem_bytes_for(int k)
{
    return (size_t)(slots_for_k(k) >> log2_group_for_k(k)) * kGroupBytes + kSmemAux;
}

__device__ __forceinline__ SmemLayout carve(unsigned char *base, uint32_t k)
{
    SmemLayout L;
    L.tab = reinterpret_cast<uint32_t *>(base);
    L.slot = reinterpret_cast<PlanSlot *>(
        base + (size_t)(slots_for_k((int)k) >> log2_group_for_k((int)k)) * kGroupBytes);
    return L;
}

// multiply four packed GF(2^8) bytes by alpha (= 2): shift left, reduce by 0x11D where bit 7 was set
__device__ __forceinline__ uint32_t xtime4(uint32_t v)
{
    const uint32_t hi = v & 0x80808080u;
    return ((v ^ hi) << 1) ^ ((hi >> 7) * 0x1du);
}

// Build the product tables for `rows` (<=4) coefficient rows coef[i*cstride + j], j < k; the
// padding tables of a partial last group are all zero.  No log/antilog lookups: entry x of
// table j is XOR_{bit b of x} (packed column j) * alpha^b, split as Hi[x >> 4] ^ Lo[x & 15].
// Caller syncs before and after.
//  * G <= 2 (k <= 14): warp-cooperative.  Each lane keeps Lo[lane & 15] / Hi[lane & 15] of its
//    sub-warp's table in two registers; an entry is two shuffles + XOR; one STS.128 instruction
//    of a warp writes 4 rows x (all replicas) = 512 B over all 32 banks in the minimum 4
//    wavefronts (a row-per-lane mapping would serialise 32-way on the replica banks).
//  * G >= 4: plain loop (generic path).
template <int NT>
__device__ __forceinline__ void build_tables(const SmemLayout &L, const uint8_t *coef, uint32_t cstride,
                                             uint32_t k, uint32_t rows)
{
    const uint32_t lg = (uint32_t)log2_group_for_k((int)k);
    const uint32_t slots = (uint32_t)slots_for_k((int)k);
    const uint32_t R = 32u >> lg;  // replicas per table = lanes per sub-warp (>= 4)
    auto packed_col = [&](uint32_t j) -> uint32_t {
        uint32_t cur = 0;
        if (j < k) {
#pragma unroll
            for (uint32_t i = 0; i < kRowsPerPass; i++)
                if (i < rows) cur |= (uint32_t)coef[i * cstride + j] << (8 * i);
        }
        return cur;
    };
    if (lg <= 1) {
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const uint32_t sub = lg ? (lane >> 4) : 0;           // table of the group this lane writes
        const uint32_t quads = R >> 2;                        // 16-byte pieces per row and table
        const uint32_t quad = lane & (quads - 1);
        const uint32_t xsel = (lane >> (lg ? 2 : 3)) & 3;     // which of the 4 rows of this instruction
        const uint32_t total = (slots >> lg) * 64;            // STS.128 warp-instructions
        const uint32_t per = (total + NT / 32 - 1) / (NT / 32);
        const uint32_t i0 = warp * per, i1 = min(total, i0 + per);
        uint32_t tprev = 0xffffffffu, lo = 0, hi = 0;
        for (uint32_t I = i0; I < i1; I++) {
            const uint32_t t = I >> 6, r4 = I & 63;
            if (t != tprev) {
                uint32_t cur = packed_col((t << lg) + sub);
                const uint32_t idx = lane & 15;
                lo = 0;
                hi = 0;
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if ((idx >> b) & 1) lo ^= cur;
                    cur = xtime4(cur);
                }
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if ((idx >> b) & 1) hi ^= cur;
                    cur = xtime4(cur);
                }
                tprev = t;
            }
            const uint32_t x = r4 * 4 + xsel;
            const uint32_t w = __shfl_sync(0xffffffffu, lo, (lane & 16) + (x & 15)) ^
                               __shfl_sync(0xffffffffu, hi, (lane & 16) + (x >> 4));
            uint32_t *dst = L.tab + (size_t)t * (kGroupBytes / 4) + x * 32 + sub * R + quad * 4;
            *reinterpret_cast<uint4 *>(dst) = make_uint4(w, w, w, w);
        }
        return;
    }
    for (uint32_t e = threadIdx.x; e < slots * 256; e += NT) {
        const uint32_t j = e >> 8, x = e & 255;
        uint32_t w = 0;
        uint32_t cur = packed_col(j);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            if ((x >> b) & 1) w ^= cur;
            cur = xtime4(cur);
        }
        // table j: group t = j >> lg, banks [sub*R, sub*R + R), row x
        uint32_t *dst = L.tab + (size_t)(j >> lg) * (kGroupBytes / 4) + x * 32 + (j & ((1u << lg) - 1)) * R;
        const uint4 v = make_uint4(w, w, w, w);
        for (uint32_t g = 0; g < R; g += 4) *reinterpret_cast<uint4 *>(dst + g) = v;
    }
}

// 16 table lookups for one 16-byte vector; acc[4*w + p] ^= T[byte p of word w]
//   base = shared address of (group, this lane's bank for this phase); row_bytes = 128
template <bool kFirst>
__device__ __forceinline__ void lookup16(uint32_t (&acc)[16], const uint4 &d, uint32_t base,
                                         uint32_t row_bytes)
{
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const uint32_t x = __byte_perm(w[i], 0, 0x4440 + p);  // byte p, zero extended (alu pipe)
            const uint32_t v = lds_u32(x * row_bytes + base);     // IMAD (fma pipe)
            if (kFirst) acc[4 * i + p] = v;
            else acc[4 * i + p] ^= v;
        }
    }
}

__device__ __forceinline__ void rows_from_acc(const uint32_t (&acc)[16], uint4 (&r)[4])
{
    uint32_t o[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
        transpose4x4(acc[4 * i + 0], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3], o[0][i], o[1][i],
                     o[2][i], o[3][i]);
#pragma unroll
    for (int row = 0; row < 4; row++) r[row] = make_uint4(o[row][0], o[row][1], o[row][2], o[row][3]);
}

// One 16-byte column of one stripe.  Slot u of a lane in sub-warp q holds source u ^ q (see file
// header); slots >= k (padding of a partial last group) hold zeros.
//   sp     : address of this column in source 0 (uniform modes) / shard 0 (plan mode)
//   kPlan  : source j lives at sp + src_off[j] (smem) instead of sp + j*stride
template <int K, bool kPlan>
__device__ __forceinline__ void column_load(const uint8_t *sp, uint32_t stride, const uint32_t *src_off,
                                            uint32_t q, uint4 (&d)[slots_for_k(K)])
{
    constexpr int S = slots_for_k(K);
#pragma unroll
    for (int u = 0; u < S; u++) {
        const uint32_t j = (uint32_t)u ^ q;
        if (S == K || j < (uint32_t)K) d[u] = ldg_stream(sp + (kPlan ? src_off[j] : j * stride));
        else d[u] = make_uint4(0, 0, 0, 0);
    }
}

// look up + transpose -> r[row] (uint4) for the slots of one column
//   tab_base : shared address of the tables
template <int K>
__device__ __forceinline__ void column_compute(uint4 (&d)[slots_for_k(K)], uint32_t tab_base, uint32_t lane,
                                               uint32_t row_bytes, uint32_t tail_bytes, uint4 (&r)[4])
{
    constexpr int S = slots_for_k(K);
    constexpr int LG = log2_group_for_k(K);
    constexpr uint32_t R = 32u >> LG;
    uint32_t acc[16];
    if (tail_bytes) {
#pragma unroll
        for (int u = 0; u < S; u++) d[u] = mask_tail(d[u], tail_bytes);
    }
#pragma unroll
    for (int u = 0; u < S; u++) {
        // group u >> LG, phase u & (G-1): this lane's bank is lane ^ (phase * R)
        const uint32_t base = tab_base + ((lane ^ (((uint32_t)u & ((1u << LG) - 1)) * R)) << 2) +
                              (uint32_t)(u >> LG) * kGroupBytes;
        if (u == 0) lookup16<true>(acc, d[u], base, row_bytes);
        else lookup16<false>(acc, d[u], base, row_bytes);
    }
    rows_from_acc(acc, r);
}

// generic k (runtime, k > 16): batches of 8 slots, the next batch prefetched while one is looked up
template <bool kPlan>
__device__ __forceinline__ void column_rows_generic(const uint8_t *sp, uint32_t stride, uint32_t k_rt,
                                                    const uint32_t *src_off, uint32_t tab_base,
                                                    uint32_t row_bytes, uint32_t tail_bytes, uint32_t lane,
                                                    uint4 (&r)[4])
{
    const uint32_t lg = (uint32_t)log2_group_for_k((int)k_rt);
    const uint32_t G = 1u << lg, R = 32u >> lg;
    const uint32_t q = lane >> (5 - lg);
    const uint32_t slots = (uint32_t)slots_for_k((int)k_rt);
    uint32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = 0;
    constexpr int W = 8;  // slots per batch; the next batch is in flight while this one is looked up
    auto load_batch = [&](uint32_t u0, uint4 (&d)[W]) {
#pragma unroll
        for (int v = 0; v < W; v++) {
            const uint32_t j = (u0 + v) ^ q;
            d[v] = make_uint4(0, 0, 0, 0);
            if (u0 + v < slots && j < k_rt) {
                d[v] = ldg_stream(sp + (kPlan ? src_off[j] : j * stride));
                if (tail_bytes) d[v] = mask_tail(d[v], tail_bytes);
            }
        }
    };
    uint4 dn[W];
    load_batch(0, dn);
    for (uint32_t u0 = 0; u0 < slots; u0 += W) {
        uint4 d[W];
#pragma unroll
        for (int v = 0; v < W; v++) d[v] = dn[v];
        if (u0 + W < slots) load_batch(u0 + W, dn);
#pragma unroll
        for (int v = 0; v < W; v++) {
            const uint32_t u = u0 + v;
            if (u < slots) {
                const uint32_t base = tab_base + ((lane ^ ((u & (G - 1)) * R)) << 2) + (u >> lg) * kGroupBytes;
                lookup16<false>(acc, d[v], base, row_bytes);
            }
        }
    }
    rows_from_acc(acc, r);
}

// ------------------------------------------------------------------ the streaming kernel
template <int K, int MODE, int NT, bool PIPE>
__global__ void __launch_bounds__(NT, GEC_MIN_BLOCKS) rs_apply_kernel(const __grid_constant__ ApplyParams p)
{
    constexpr int kThreads = NT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const SmemLayout L = carve(smem_raw, p.k);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = (K > 0) ? (uint32_t)K : p.k;
    constexpr int KD = K > 0 ? K : 1;
    constexpr int SD = slots_for_k(KD);

    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(L.tab);
    const uint32_t q = lane >> (5 - log2_group_for_k((int)k));  // sub-warp index

    if (MODE != kModePlan) {
        // ---- uniform coefficient matrix: build once, then warps stream items ----------------
        build_tables<NT>(L, p.coef, k, k, p.rows);
        __syncthreads();

        const uint32_t ips = p.items_per_stripe;
        const uint32_t total = p.n * ips;  // host guarantees < 2^32
        const uint32_t gwarps = gridDim.x * (kThreads / 32);

        // position of this lane's column for a work item (stripe, 32-column chunk)
        struct Pos {
            const uint8_t *sp;
            uint32_t s, col, tail;
            bool valid;
        };
        auto locate = [&](uint32_t item) -> Pos {
            Pos z;
            z.s = item / ips;
            const uint32_t c = item - z.s * ips;
            const uint32_t len = p.shard_len ? __ldg(p.shard_len + z.s) : p.stride;
            const uint32_t nvec = (len + 15) >> 4;
            z.col = c * 32 + lane;
            z.valid = z.col < nvec;
            z.tail = (z.col == nvec - 1) ? (len & 15) : 0;
            z.sp = p.src + (unsigned long long)z.s * p.src_pitch + (size_t)z.col * 16;
            return z;
        };
        auto finish = [&](const Pos &z, const uint4 (&r)[4]) {
            uint32_t mm = 0;
            if (MODE == kModeEncode) {
                if (z.valid) {
                    uint8_t *dp = p.dst + (unsigned long long)z.s * p.dst_pitch + (size_t)z.col * 16;
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (i < (int)p.rows) stg_stream(dp + (size_t)i * p.stride, r[i]);
                }
            } else {
                if (z.valid) {
                    // stored parity rows follow the k data shards of the same stripe
                    const uint8_t *pp = z.sp + (size_t)(k + p.row_off) * p.stride;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (i < (int)p.rows) {
                            uint4 st = ldg_stream(pp + (size_t)i * p.stride);
                            if (z.tail) st = mask_tail(st, z.tail);
                            const uint32_t diff = (st.x ^ r[i].x) | (st.y ^ r[i].y) | (st.z ^ r[i].z) |
                                                  (st.w ^ r[i].w);
                            if (diff) mm |= 1u << (p.row_off + i);
                        }
                    }
                }
                // warp-shuffle OR reduction of the per-lane mismatch bits, one atomic per warp
                mm = __reduce_or_sync(0xffffffffu, mm);
                if (mm && lane == 0) atomicOr(p.mismatch + z.s, mm);
            }
        };

        uint32_t item = blockIdx.x * (kThreads / 32) + warp;
        if (K > 0 && PIPE) {
            uint4 dn[SD];
            Pos nx;
            nx.valid = false;
            if (item < total) {
                nx = locate(item);
                if (nx.valid) column_load<KD, false>(nx.sp, p.stride, nullptr, q, dn);
            }
            while (item < total) {
                uint4 d[SD];
#pragma unroll
                for (int u = 0; u < SD; u++) d[u] = dn[u];
                const Pos cur = nx;
                item += gwarps;
                nx.valid = false;
                if (item < total) {
                    nx = locate(item);
                    if (nx.valid) column_load<KD, false>(nx.sp, p.stride, nullptr, q, dn);
                }
                uint4 r[4];
                if (cur.valid) column_compute<KD>(d, tab_base, lane, p.row_bytes, cur.tail, r);
                finish(cur, r);
            }
        } else {
            for (; item < total; item += gwarps) {
                const Pos cur = locate(item);
                uint4 r[4];
                if (cur.valid) {
                    if (K > 0) {
                        uint4 d[SD];
                        column_load<KD, false>(cur.sp, p.stride, nullptr, q, d);
                        column_compute<KD>(d, tab_base, lane, p.row_bytes, cur.tail, r);
                    } else {
                        column_rows_generic<false>(cur.sp, p.stride, k, nullptr, tab_base, p.row_bytes,
                                                   cur.tail, lane, r);
                    }
                }
                finish(cur, r);
            }
        }
    } else {
        // ---- per-stripe matrices: CTAs pull stripes from an atomic counter -------------------
        // Two plan slots: while the CTA streams stripe `cur`, warp 0 claims the next stripe and
        // stages its plan; at the stripe boundary the first column loads of the new stripe are
        // issued BEFORE its tables are rebuilt, so HBM latency hides behind the rebuild.
        // Stripes are claimed one staging ahead: the atomic issued now is consumed by the NEXT call,
        // so its round trip never sits between the CTA and the barrier that waits for the plan.
        uint32_t claimed = 0;                  // meaningful in lane 0 of warp 0
        auto stage_plan = [&](PlanSlot &ps) {  // executed by warp 0 only
            const uint32_t s = __shfl_sync(0xffffffffu, claimed, 0);
            if (lane == 0) claimed = atomicAdd(p.counter, 1u);
            int rows = 0;
            if (s < p.n) {
                const StripePlan *pl = p.plan + s;
                rows = pl->unrecoverable ? 0 : min(kRowsPerPass, (int)pl->nrows - (int)p.row_off);
                if (rows > 0) {
                    if (lane < k) ps.src_off[lane] = (uint32_t)pl->surv[lane] * p.stride;
                    if (lane < (uint32_t)rows) ps.dst_off[lane] = (uint32_t)pl->out_idx[p.row_off + lane] * p.stride;
                    for (uint32_t e = lane; e < (uint32_t)rows * kMaxK; e += 32)
                        ps.coef[e] = pl->coef[p.row_off + e / kMaxK][e % kMaxK];
                    if (lane == 0) {
                        ps.key_present = pl->key_present;
                        ps.key_out = pl->key_out;
                    }
                }
            }
            if (lane == 0) {
                ps.sid = s;
                ps.rows = rows;
                ps.chunk_next = kThreads / 32;  // chunk w < #warps belongs to warp w, the rest are claimed
            }
        };
        if (warp == 0) {
            if (lane == 0) claimed = atomicAdd(p.counter, 1u);
            stage_plan(L.slot[0]);
        }
        __syncthreads();
        unsigned long long built_present = ~0ull, built_out = ~0ull;
        bool have_tables = false;
        bool have_dn = false;  // dn already holds this thread's first column of the stripe in slot cs
        uint4 dn[SD];
        for (uint32_t cs = 0;; cs ^= 1) {
            const PlanSlot &ps = L.slot[cs];
            const uint32_t s = ps.sid;
            if (s >= p.n) break;
            const int rows = ps.rows;
            uint32_t len = 0, nvec = 0;
            const uint8_t *sbase = p.src + (unsigned long long)s * p.src_pitch;
            uint8_t *dbase = p.dst + (unsigned long long)s * p.dst_pitch;
            uint32_t col = tid;
            if (rows > 0) {
                len = p.shard_len ? __ldg(p.shard_len + s) : p.stride;
                nvec = (len + 15) >> 4;
                if (K > 0 && PIPE && col < nvec && !have_dn)
                    column_load<KD, true>(sbase + (size_t)col * 16, p.stride, ps.src_off, q, dn);
            }
            have_dn = false;
            if (warp == 0) stage_plan(L.slot[cs ^ 1]);
            if (rows > 0 && !(have_tables && ps.key_present == built_present && ps.key_out == built_out)) {
                build_tables<NT>(L, ps.coef, kMaxK, k, (uint32_t)rows);
                built_present = ps.key_present;
                built_out = ps.key_out;
                have_tables = true;
            }
            __syncthreads();  // tables of `s` complete, next plan staged
            if (rows > 0) {
                if (K > 0 && PIPE) {
                    // 32-column chunks: chunk `warp` is this warp's first one, the rest are claimed
                    // from a shared counter one iteration ahead, so every warp reaches the barrier
                    // within one chunk of the others (a static col = tid + i*NT split lets the slow
                    // warps' lag accumulate over the ~13 iterations of a stripe)
                    // claims run one iteration ahead of their use, so the shared-memory atomic
                    // never delays the prefetch loads.  Measured (profiles/r01_summary.md): RS(10,4)
                    // with 4 erasures 0.905 -> 0.926, RS(6,3) 0.867 -> 0.905, RS(4,2) 0.949 -> 0.976;
                    // a single erasure per stripe is slower than with the static split (0.91 -> 0.86).
                    uint32_t chunk = warp, nxt = 0, nxt2 = 0;
                    if (lane == 0) nxt = atomicAdd(&L.slot[cs].chunk_next, 1u);
                    nxt = __shfl_sync(0xffffffffu, nxt, 0);
                    while (chunk * 32 < nvec) {
                        uint4 d[SD];
#pragma unroll
                        for (int u = 0; u < SD; u++) d[u] = dn[u];
                        const uint32_t cur = chunk * 32 + lane;
                        if (lane == 0) nxt2 = atomicAdd(&L.slot[cs].chunk_next, 1u);  // consumed next iteration
                        if (nxt * 32 < nvec) {
                            const uint32_t ncol = nxt * 32 + lane;
                            if (ncol < nvec)
                                column_load<KD, true>(sbase + (size_t)ncol * 16, p.stride, ps.src_off, q, dn);
                        } else {
                            // no chunk left in this stripe for this warp: start on the next stripe
                            // (its plan was staged before the barrier above) so HBM never drains
                            const PlanSlot &nx = L.slot[cs ^ 1];
                            if (nx.sid < p.n && nx.rows > 0) {
                                const uint32_t nlen = p.shard_len ? __ldg(p.shard_len + nx.sid) : p.stride;
                                if (tid < ((nlen + 15) >> 4)) {
                                    column_load<KD, true>(p.src + (unsigned long long)nx.sid * p.src_pitch +
                                                              (size_t)tid * 16,
                                                          p.stride, nx.src_off, q, dn);
                                    have_dn = true;
                                }
                            }
                        }
                        if (cur < nvec) {
                            const uint32_t tail = (cur == nvec - 1) ? (len & 15) : 0;
                            uint4 r[4];
                            column_compute<KD>(d, tab_base, lane, p.row_bytes, tail, r);
#pragma unroll
                            for (int i = 0; i < 4; i++)
                                if (i < rows) stg_stream(dbase + ps.dst_off[i] + (size_t)cur * 16, r[i]);
                        }
                        chunk = nxt;
                        nxt = __shfl_sync(0xffffffffu, nxt2, 0);
                    }
                } else {
                    for (; col < nvec; col += kThreads) {
                        const uint32_t tail = (col == nvec - 1) ? (len & 15) : 0;
                        uint4 r[4];
                        if (K > 0) {
                            uint4 d[SD];
                            column_load<KD, true>(sbase + (size_t)col * 16, p.stride, ps.src_off, q, d);
                            column_compute<KD>(d, tab_base, lane, p.row_bytes, tail, r);
                        } else {
                            column_rows_generic<true>(sbase + (size_t)col * 16, p.stride, k, ps.src_off, tab_base,
                                                      p.row_bytes, tail, lane, r);
                        }
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (i < rows) stg_stream(dbase + ps.dst_off[i] + (size_t)col * 16, r[i]);
                    }
                }
            }
            __syncthreads();  // tables and slot `cs` are free again
        }
    }
}

// ------------------------------------------------------------------ decode planning
struct PlanParams {
    const uint8_t *present;  // n*(k+m)
    const uint8_t *want;     // nullable
    int32_t *status;         // nullable, n
    StripePlan *plan;        // n
    uint32_t *counter;       // reset to 0 for the apply kernel's scheduler
    uint32_t n, k, m;
    uint32_t present_is_bad;  // 1: `present` holds "bad" flags (scrub): a shard is present iff flag == 0
    alignas(16) uint8_t P[kMaxM * kMaxK];  // parity rows, P[i*k + j] (16-byte aligned: staged with uint4 loads)
};

constexpr int kPlanWarps = 4;

// One warp per stripe: pick the first k present shards, solve the a x a system that couples the
// a absent data shards to the a parity survivors (Gauss-Jordan, lanes own columns), compose the
// rows that map the survivors straight to every wanted absent shard.
__global__ void __launch_bounds__(kPlanWarps * 32) rs_plan_kernel(const __grid_constant__ PlanParams q)
{
    __shared__ __align__(16) uint8_t s_gf[sizeof(GfTables)];
    uint8_t *const s_exp = s_gf, *const s_log = s_gf + 512;
    __shared__ uint8_t s_A[kPlanWarps][kMaxK][2 * kMaxK];
    __shared__ uint8_t s_surv[kPlanWarps][kMaxK];
    __shared__ uint8_t s_out[kPlanWarps][kMaxM];
    __shared__ uint8_t s_col[kPlanWarps][kMaxK];
    __shared__ uint8_t s_dm[kPlanWarps][kMaxM];
    __shared__ __align__(16) uint8_t s_P[kMaxM * kMaxK];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_gf_tables(s_gf);
    {  // parity rows: kernel parameter (constant bank) -> shared, warp-uniform 16-byte loads
        const uint4 *src = reinterpret_cast<const uint4 *>(q.P);
        for (uint32_t i = warp; i < sizeof(q.P) / 16; i += blockDim.x >> 5) {
            const uint4 v = src[i];
            if (lane < 4) reinterpret_cast<uint32_t *>(s_P)[i * 4 + lane] = lane == 0 ? v.x : (lane == 1 ? v.y : (lane == 2 ? v.z : v.w));
        }
    }
    if (blockIdx.x == 0 && tid == 0) *q.counter = 0;
    __syncthreads();
    auto mul = [&](uint32_t a, uint32_t b) -> uint32_t {
        return (a && b) ? s_exp[s_log[a] + s_log[b]] : 0u;
    };

    const uint32_t s = blockIdx.x * kPlanWarps + warp;
    if (s >= q.n) return;
    const uint32_t k = q.k, m = q.m, tot = k + m;
    StripePlan *pl = q.plan + s;
    const uint8_t *pr = q.present + (size_t)s * tot;
    const uint8_t *wn = q.want ? q.want + (size_t)s * tot : nullptr;

    // presence / wanted masks (tot <= 40: two ballot rounds)
    const bool inv = q.present_is_bad != 0;
    const bool p0 = lane < tot && ((pr[lane] != 0) != inv);
    const bool p1 = lane + 32 < tot && ((pr[lane + 32] != 0) != inv);
    const unsigned long long present =
        (unsigned long long)__ballot_sync(0xffffffffu, p0) |
        ((unsigned long long)__ballot_sync(0xffffffffu, p1) << 32);
    const bool w0 = lane < tot && !p0 && (!wn || wn[lane] != 0);
    const bool w1 = lane + 32 < tot && !p1 && (!wn || wn[lane + 32] != 0);
    unsigned long long outmask = (unsigned long long)__ballot_sync(0xffffffffu, w0) |
                                 ((unsigned long long)__ballot_sync(0xffffffffu, w1) << 32);
    const int npresent = __popcll(present);
    const bool bad = npresent < (int)k;
    if (bad) outmask = 0;
    int nrows = __popcll(outmask);

    if (lane == 0) {
        // survivors: first k present; outputs: wanted absent shards in index order
        unsigned long long pm = present;
        for (uint32_t r = 0; r < k && pm; r++) {
            const int i = __ffsll((long long)pm) - 1;
            s_surv[warp][r] = (uint8_t)i;
            pm &= pm - 1;
        }
        unsigned long long om = outmask;
        for (int r = 0; r < nrows; r++) {
            const int i = __ffsll((long long)om) - 1;
            s_out[warp][r] = (uint8_t)i;
            om &= om - 1;
        }
        pl->key_present = present;
        pl->key_out = outmask;
        pl->nrows = (uint8_t)nrows;
        pl->unrecoverable = bad ? 1 : 0;
        if (q.status) q.status[s] = bad ? -4 : 0;
    }
    __syncwarp();
    if (bad || nrows == 0) return;

    // The survivors are the present data shards (identity rows of the generator) followed by the
    // first `a` present parity shards, a = number of absent data shards.  Only an a x a system
    // couples the unknowns:   y_Pu = A x_Dm + B x_Dp,   A = P[Pu][Dm], B = P[Pu][Dp]
    //   =>  x_Dm = Ainv y_Pu + (Ainv B) x_Dp          (characteristic 2)
    // so a full k x k inversion is never needed (a <= m <= 8).
    uint8_t(*A)[2 * kMaxK] = s_A[warp];  // rows 0..7: [A | I] -> [I | Ainv]; rows 8..15: Ainv*B; rows 16..23: decode rows
    const unsigned long long datamask = (k >= 64 ? ~0ull : ((1ull << k) - 1));
    const unsigned long long mdmask = ~present & datamask;
    const uint32_t a = (uint32_t)__popcll(mdmask);
    const uint32_t nd = k - a;  // data survivors occupy positions [0, nd) of surv[], parity survivors [nd, k)
    if (lane == 0) {
        unsigned long long mm = mdmask;
        for (uint32_t t = 0; t < a; t++) {
            s_dm[warp][t] = (uint8_t)(__ffsll((long long)mm) - 1);
            mm &= mm - 1;
        }
    }
    __syncwarp();
    if (a) {
        for (uint32_t r = 0; r < a; r++) {
            const uint32_t prow = (uint32_t)s_surv[warp][nd + r] - k;
            if (lane < 2 * a) A[r][lane] = lane < a ? s_P[prow * k + s_dm[warp][lane]] : (uint8_t)(lane - a == r);
        }
        __syncwarp();
        for (uint32_t c = 0; c < a; c++) {
            const bool cand = lane < a && lane >= c && A[lane][c] != 0;  // pivot search: lanes are rows
            const uint32_t bal = __ballot_sync(0xffffffffu, cand);
            if (!bal) {  // singular: cannot happen for an MDS generator; report unrecoverable
                if (lane == 0) {
                    pl->nrows = 0;
                    pl->unrecoverable = 1;
                    if (q.status) q.status[s] = -4;
                }
                return;
            }
            const uint32_t piv = __ffs(bal) - 1;
            const uint32_t iv = s_exp[255 - s_log[A[piv][c]]];
            __syncwarp();
            if (lane < 2 * a) {
                const uint8_t x = A[piv][lane], y = A[c][lane];
                A[piv][lane] = y;                  // swap (no-op when piv == c)
                A[c][lane] = (uint8_t)mul(x, iv);  // scaled pivot row
            }
            __syncwarp();
            const uint32_t pv = lane < 2 * a ? A[c][lane] : 0;
            if (lane < a) s_col[warp][lane] = lane == c ? 0 : A[lane][c];
            __syncwarp();
            for (uint32_t r = 0; r < a; r++) {
                const uint32_t f = s_col[warp][r];  // broadcast; lanes own columns
                if (f && lane < 2 * a) A[r][lane] ^= (uint8_t)mul(f, pv);
            }
            __syncwarp();
        }
        // decode row of missing data shard dm[t] over the survivor positions:
        //   data survivors:   (Ainv B)[t][j] = XOR_r Ainv[t][r] * P[Pu_r][surv[j]]
        //   parity survivors: Ainv[t][j - nd]
        for (uint32_t t = 0; t < a; t++) {
            uint32_t v = 0;
            if (lane < nd) {
                const uint32_t col = s_surv[warp][lane];
                for (uint32_t r = 0; r < a; r++)
                    v ^= mul(A[t][a + r], s_P[((uint32_t)s_surv[warp][nd + r] - k) * k + col]);
            } else if (lane < k) {
                v = A[t][a + (lane - nd)];
            }
            if (lane < k) A[16 + t][lane] = (uint8_t)v;
        }
        __syncwarp();
    }
    // rows for the wanted outputs
    for (int r = 0; r < nrows; r++) {
        const uint32_t o = s_out[warp][r];
        if (lane < k) {
            uint32_t v = 0;
            if (o < k) {
                // rank of o among the absent data shards
                const uint32_t t = (uint32_t)__popcll(mdmask & ((1ull << o) - 1));
                v = A[16 + t][lane];
            } else {
                // absent parity row i: P[i][Dp] on the data survivors, plus P[i][Dm] through the rows above
                const uint32_t i = o - k;
                if (lane < nd) v = s_P[i * k + s_surv[warp][lane]];
                for (uint32_t t = 0; t < a; t++) v ^= mul(s_P[i * k + s_dm[warp][t]], A[16 + t][lane]);
            }
            pl->coef[r][lane] = (uint8_t)v;
        }
    }
    if (lane < k) pl->surv[lane] = s_surv[warp][lane];
    if (lane < (uint32_t)nrows) pl->out_idx[lane] = s_out[warp][lane];
}

// ------------------------------------------------------------------ synthetic data
// splitmix64 counter stream of SURVEY.md 8(d) (the CPU checker generates the same stream).
__global__ void fill_random_kernel(unsigned long long *dst, size_t nwords, unsigned long long seed,
                                   unsigned long long first_idx)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords;
         i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = seed + (first_idx + i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        dst[i] = z ^ (z >> 31);
    }
}

// ------------------------------------------------------------------ per-shard integrity (row f2)
// blake2sum (BLAKE2b-512 truncated to 32 bytes, src/util/data.rs:130-138) of every shard of a
// batch, one thread per shard.  BLAKE2b is sequential per message, so the parallelism is across
// the n * shards_per_stripe independent shards; compute-bound (~28 integer instructions per
// byte), not HBM-bound.  `expect` != NULL turns it into the scrub check: bad[i] = sum differs.
struct SumParams {
    const uint8_t *base;        // shard i at base + i * stride
    const uint32_t *shard_len;  // per stripe, nullable (=> stride)
    const uint8_t *expect;      // nullable: 32 bytes per shard to compare with
    uint8_t *sums;              // nullable: 32 bytes per shard out
    uint8_t *bad;               // nullable: 1 byte per shard out (only with expect)
    uint32_t stride;
    uint32_t per_stripe;        // shards per stripe (k, m or k+m)
    uint32_t n_shards;
    uint32_t out_per_stripe;    // sums/expect/bad are indexed (i / per_stripe) * out_per_stripe
    uint32_t out_off;           //                              + out_off + i % per_stripe
    // optional second segment (shards n_first.. of the launch): lets one launch hash the data
    // array AND the parity array of an encode batch (half the latency of two launches)
    const uint8_t *base2;
    uint32_t n_first;           // shards in the first segment (== n_shards when there is no second)
    uint32_t per_stripe2;
    uint32_t out_off2;
};

// where shard i of a launch lives, which stripe's length applies, and its slot in sums/expect/bad
__device__ __forceinline__ void locate_shard(const SumParams &q, uint32_t i, const uint8_t *&p, uint32_t &len,
                                             size_t &oi)
{
    uint32_t stripe;
    if (i < q.n_first) {
        stripe = i / q.per_stripe;
        p = q.base + (size_t)i * q.stride;
        oi = (size_t)stripe * q.out_per_stripe + q.out_off + (i - stripe * q.per_stripe);
    } else {
        const uint32_t j = i - q.n_first;
        stripe = j / q.per_stripe2;
        p = q.base2 + (size_t)j * q.stride;
        oi = (size_t)stripe * q.out_per_stripe + q.out_off2 + (j - stripe * q.per_stripe2);
    }
    len = q.shard_len ? __ldg(q.shard_len + stripe) : q.stride;
}

__global__ void __launch_bounds__(128) blake2sum_shards_kernel(const __grid_constant__ SumParams q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= q.n_shards) return;
    const uint8_t *p;
    uint32_t len;
    size_t oi;
    locate_shard(q, i, p, len, oi);
    Blake2bState S;
    blake2b_init512(S);
    uint64_t m[16];
    uint32_t off = 0;
    while (len - off > 128) {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + off + 16 * w);
            m[2 * w] = (uint64_t)v.x | ((uint64_t)v.y << 32);
            m[2 * w + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
        }
        S.t += 128;
        blake2b_compress(S, m, false);
        off += 128;
    }
    const uint32_t rem = len - off;  // 0..128 bytes in the last block (0 only for an empty shard)
#pragma unroll
    for (int w = 0; w < 8; w++) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((uint32_t)(16 * w) < rem) {
            v = *reinterpret_cast<const uint4 *>(p + off + 16 * w);  // within roundup16(len): readable
            if (rem - 16 * w < 16) v = mask_tail(v, rem - 16 * w);
        }
        m[2 * w] = (uint64_t)v.x | ((uint64_t)v.y << 32);
        m[2 * w + 1] = (uint64_t)v.z | ((uint64_t)v.w << 32);
    }
    S.t += rem;
    blake2b_compress(S, m, true);
    if (q.sums) {
        uint4 *o = reinterpret_cast<uint4 *>(q.sums + oi * 32);
        o[0] = make_uint4((uint32_t)S.h[0], (uint32_t)(S.h[0] >> 32), (uint32_t)S.h[1], (uint32_t)(S.h[1] >> 32));
        o[1] = make_uint4((uint32_t)S.h[2], (uint32_t)(S.h[2] >> 32), (uint32_t)S.h[3], (uint32_t)(S.h[3] >> 32));
    }
    if (q.expect && q.bad) {
        const unsigned long long *e = reinterpret_cast<const unsigned long long *>(q.expect + oi * 32);
        q.bad[oi] = (e[0] != S.h[0]) | (e[1] != S.h[1]) | (e[2] != S.h[2]) | (e[3] != S.h[3]);
    }
}


// ---- the same hash with FOUR lanes per shard --------------------------------------------------
// With a few ten-thousand shards per batch the one-thread-per-shard kernel leaves most of the GPU
// idle (28 672 shards = 1.5 warps per scheduler).  BLAKE2b's four column G functions and four
// diagonal G functions are independent, so a quad of lanes shares one message: lane c owns
// column c of the 4x4 state (a,b,c,d = v[c], v[4+c], v[8+c], v[12+c]); the diagonal step is the
// column step after rotating rows 1,2,3 by 1,2,3 lanes inside the quad (quad-masked shuffles).
// The 128-byte message block sits in shared memory (32 B loaded per lane, coalesced per quad); the
// per-round word selection sigma[r] is four byte offsets packed in one register per round.
// ~1.5x the instructions of the scalar kernel for 4x the parallelism.
__constant__ uint8_t c_b2_sigma[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

__device__ __forceinline__ uint64_t shfl64(uint32_t mask, uint64_t v, uint32_t src)
{
    const uint32_t lo = __shfl_sync(mask, (uint32_t)v, src), hi = __shfl_sync(mask, (uint32_t)(v >> 32), src);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
__device__ __forceinline__ uint64_t lds_u64(uint32_t addr)
{
    uint64_t r;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(addr));
    return r;
}

constexpr int kQuadThreads = 128;  // 32 shards per block

__global__ void __launch_bounds__(kQuadThreads) blake2sum_shards_quad_kernel(const __grid_constant__ SumParams q)
{
    // 144-byte pitch: consecutive quads start 4 banks apart (a 128-byte pitch puts the same word
    // of all 8 quads of a warp in the same bank: 8-way conflicts on every message load)
    __shared__ __align__(16) uint8_t s_msg[kQuadThreads / 4][144];
    const uint32_t tid = threadIdx.x, lane = tid & 31, c = tid & 3;
    const uint32_t i = blockIdx.x * (kQuadThreads / 4) + (tid >> 2);  // shard index
    if (i >= q.n_shards) return;                                      // whole quads leave together
    const uint32_t qmask = 0xFu << (lane & ~3u), qbase = lane & ~3u;
    const uint8_t *p;
    uint32_t len;
    size_t oi;
    locate_shard(q, i, p, len, oi);
    const uint32_t msg = (uint32_t)__cvta_generic_to_shared(&s_msg[tid >> 2][0]);

    // per-round message offsets for this lane: {col x, col y, diag x, diag y} * 8 bytes
    uint32_t off[12];
#pragma unroll
    for (int r = 0; r < 12; r++)
        off[r] = ((uint32_t)c_b2_sigma[r][2 * c] << 3) | ((uint32_t)c_b2_sigma[r][2 * c + 1] << 11) |
                 ((uint32_t)c_b2_sigma[r][8 + 2 * c] << 19) | ((uint32_t)c_b2_sigma[r][8 + 2 * c + 1] << 27);
    const uint64_t IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                            0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
    const uint64_t ivlo = c == 0 ? IV[0] : (c == 1 ? IV[1] : (c == 2 ? IV[2] : IV[3]));
    const uint64_t ivhi = c == 0 ? IV[4] : (c == 1 ? IV[5] : (c == 2 ? IV[6] : IV[7]));
    uint64_t h0 = ivlo ^ (c == 0 ? 0x01010040ull : 0ull), h1 = ivhi;  // h[c], h[4+c]

    // this lane's 32 bytes of the block starting at byte `o` (zero beyond len)
    auto load32 = [&](uint32_t o, uint4 &x, uint4 &y) {
        x = make_uint4(0, 0, 0, 0);
        y = make_uint4(0, 0, 0, 0);
        const uint32_t a0 = o + 32 * c;
        if (a0 < len) {
            x = *reinterpret_cast<const uint4 *>(p + a0);
            if (len - a0 < 16) x = mask_tail(x, len - a0);
        }
        if (a0 + 16 < len) {
            y = *reinterpret_cast<const uint4 *>(p + a0 + 16);
            if (len - a0 - 16 < 16) y = mask_tail(y, len - a0 - 16);
        }
    };
    uint4 nx, ny;
    load32(0, nx, ny);
    uint32_t o = 0;
    for (;;) {
        const bool last = len - o <= 128;  // also true for an empty shard
        __syncwarp(qmask);                  // previous block's words are no longer needed
        *reinterpret_cast<uint4 *>(&s_msg[tid >> 2][32 * c]) = nx;
        *reinterpret_cast<uint4 *>(&s_msg[tid >> 2][32 * c + 16]) = ny;
        __syncwarp(qmask);
        if (!last) load32(o + 128, nx, ny);  // prefetch the next block
        const uint64_t t = last ? (uint64_t)len : (uint64_t)o + 128;
        uint64_t va = h0, vb = h1, vc = ivlo, vd = ivhi;
        if (c == 0) vd ^= t;
        if (c == 2 && last) vd = ~vd;
#pragma unroll
        for (int r = 0; r < 12; r++) {
            const uint32_t f = off[r];
            uint64_t mx = lds_u64(msg + (f & 0xff)), my = lds_u64(msg + ((f >> 8) & 0xff));
            GEC_B2_G(va, vb, vc, vd, mx, my);
            vb = shfl64(qmask, vb, qbase + ((c + 1) & 3));
            vc = shfl64(qmask, vc, qbase + ((c + 2) & 3));
            vd = shfl64(qmask, vd, qbase + ((c + 3) & 3));
            mx = lds_u64(msg + ((f >> 16) & 0xff));
            my = lds_u64(msg + (f >> 24));
            GEC_B2_G(va, vb, vc, vd, mx, my);
            vb = shfl64(qmask, vb, qbase + ((c + 3) & 3));
            vc = shfl64(qmask, vc, qbase + ((c + 2) & 3));
            vd = shfl64(qmask, vd, qbase + ((c + 1) & 3));
        }
        h0 ^= va ^ vc;
        h1 ^= vb ^ vd;
        if (last) break;
        o += 128;
    }
    if (q.sums) *reinterpret_cast<unsigned long long *>(q.sums + oi * 32 + 8 * c) = h0;
    if (q.expect && q.bad) {
        const unsigned long long e = *reinterpret_cast<const unsigned long long *>(q.expect + oi * 32 + 8 * c);
        const uint32_t diff = __ballot_sync(qmask, e != h0) & qmask;
        if (c == 0) q.bad[oi] = diff ? 1 : 0;
    }
}

}  // namespace garage_ec

/*
 * rs_oracle.c -- scalar CPU oracle (normative definition; see rs_oracle.h header).
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED by the reference (it has no RS code);
 * pinned by SURVEY.md section 8(c) known-answer vectors + oracle/rs_oracle_np.py.
 *
 * One byte at a time, log/exp tables; deliberately the dumbest possible statement of
 * the arithmetic so that it can be audited by eye.
 */
#include "rs_oracle.h"

#include <stdlib.h>
#include <string.h>

static uint8_t g_exp[512];
static uint8_t g_log[256];
static int g_init = 0;

static void init_tables(void)
{
	if (g_init) return;
	unsigned x = 1;
	for (int i = 0; i < 255; i++) {
		g_exp[i] = (uint8_t)x;
		g_log[x] = (uint8_t)i;
		x <<= 1;
		if (x & 0x100) x ^= 0x11D; /* x^8 + x^4 + x^3 + x^2 + 1 */
	}
	for (int i = 255; i < 512; i++) g_exp[i] = g_exp[i - 255];
	g_log[0] = 0; /* undefined; never used */
	g_init = 1;
}

uint8_t rs_oracle_gf_mul(uint8_t a, uint8_t b)
{
	init_tables();
	if (a == 0 || b == 0) return 0;
	return g_exp[g_log[a] + g_log[b]];
}

uint8_t rs_oracle_gf_inv(uint8_t a)
{
	init_tables();
	return g_exp[255 - g_log[a]];
}

uint8_t rs_oracle_gf_exp(int i)
{
	init_tables();
	return g_exp[i];
}

uint8_t rs_oracle_gf_log(uint8_t a)
{
	init_tables();
	return g_log[a];
}

static uint8_t gf_pow(uint8_t a, int n)
{
	/* 0^0 = 1 by convention */
	uint8_t r = 1;
	for (int i = 0; i < n; i++) r = rs_oracle_gf_mul(r, a);
	return r;
}

int rs_oracle_invert(uint8_t *M, int n)
{
	init_tables();
	uint8_t *w = (uint8_t *)malloc((size_t)n * 2 * n);
	if (!w) return -1;
	const int W = 2 * n;
	for (int r = 0; r < n; r++) {
		memcpy(w + r * W, M + r * n, n);
		memset(w + r * W + n, 0, n);
		w[r * W + n + r] = 1;
	}
	for (int c = 0; c < n; c++) {
		int piv = -1;
		for (int r = c; r < n; r++)
			if (w[r * W + c]) { piv = r; break; }
		if (piv < 0) { free(w); return -1; }
		if (piv != c)
			for (int x = 0; x < W; x++) {
				uint8_t t = w[c * W + x];
				w[c * W + x] = w[piv * W + x];
				w[piv * W + x] = t;
			}
		uint8_t inv = rs_oracle_gf_inv(w[c * W + c]);
		for (int x = 0; x < W; x++) w[c * W + x] = rs_oracle_gf_mul(w[c * W + x], inv);
		for (int r = 0; r < n; r++) {
			if (r == c) continue;
			uint8_t f = w[r * W + c];
			if (!f) continue;
			for (int x = 0; x < W; x++) w[r * W + x] ^= rs_oracle_gf_mul(f, w[c * W + x]);
		}
	}
	for (int r = 0; r < n; r++) memcpy(M + r * n, w + r * W + n, n);
	free(w);
	return 0;
}

int rs_oracle_build_matrix(int k, int m, int kind, uint8_t *P)
{
	init_tables();
	if (k < 1 || m < 1 || k + m > 256) return -1;
	if (kind == RS_ORACLE_CAUCHY) {
		for (int i = 0; i < m; i++)
			for (int j = 0; j < k; j++)
				P[i * k + j] = rs_oracle_gf_inv((uint8_t)((k + i) ^ j));
		return 0;
	}
	if (kind != RS_ORACLE_VANDERMONDE) return -1;
	const int n = k + m;
	uint8_t *V = (uint8_t *)malloc((size_t)n * k);
	uint8_t *T = (uint8_t *)malloc((size_t)k * k);
	if (!V || !T) { free(V); free(T); return -1; }
	for (int r = 0; r < n; r++)
		for (int c = 0; c < k; c++) V[r * k + c] = gf_pow((uint8_t)r, c);
	memcpy(T, V, (size_t)k * k);
	if (rs_oracle_invert(T, k)) { free(V); free(T); return -1; }
	for (int i = 0; i < m; i++)
		for (int j = 0; j < k; j++) {
			uint8_t acc = 0;
			for (int x = 0; x < k; x++)
				acc ^= rs_oracle_gf_mul(V[(k + i) * k + x], T[x * k + j]);
			P[i * k + j] = acc;
		}
	free(V);
	free(T);
	return 0;
}

static void encode_rows(int k, int rows, const uint8_t *C /* rows x k */,
			const uint8_t *const *in, uint8_t *const *out, size_t len)
{
	for (int i = 0; i < rows; i++) {
		uint8_t *o = out[i];
		for (size_t t = 0; t < len; t++) {
			uint8_t acc = 0;
			for (int j = 0; j < k; j++) acc ^= rs_oracle_gf_mul(C[i * k + j], in[j][t]);
			o[t] = acc;
		}
	}
}

void rs_oracle_encode(int k, int m, const uint8_t *P, const uint8_t *data, uint8_t *parity,
		      const uint32_t *shard_len, size_t stride, size_t n)
{
	init_tables();
	const uint8_t *in[256];
	uint8_t *out[256];
	for (size_t s = 0; s < n; s++) {
		size_t len = shard_len ? shard_len[s] : stride;
		for (int j = 0; j < k; j++) in[j] = data + (s * k + j) * stride;
		for (int i = 0; i < m; i++) out[i] = parity + (s * m + i) * stride;
		encode_rows(k, m, P, in, out, len);
	}
}

size_t rs_oracle_reconstruct(int k, int m, const uint8_t *P, uint8_t *shards,
			     const uint8_t *present, int32_t *status, const uint32_t *shard_len,
			     size_t stride, size_t n)
{
	init_tables();
	const int tot = k + m;
	size_t bad = 0;
	uint8_t *sub = (uint8_t *)malloc((size_t)k * k);
	uint8_t *rows = (uint8_t *)malloc((size_t)tot * k);
	for (size_t s = 0; s < n; s++) {
		size_t len = shard_len ? shard_len[s] : stride;
		const uint8_t *pr = present + s * tot;
		uint8_t *base = shards + s * tot * stride;
		int surv[256], ns = 0, npresent = 0;
		for (int i = 0; i < tot; i++) npresent += pr[i] ? 1 : 0;
		if (npresent < k) {
			if (status) status[s] = -1;
			bad++;
			continue;
		}
		if (status) status[s] = 0;
		if (npresent == tot) continue;
		/* first k present rows of the generator [I ; P] */
		for (int i = 0; i < tot && ns < k; i++)
			if (pr[i]) surv[ns++] = i;
		for (int r = 0; r < k; r++) {
			int g = surv[r];
			for (int c = 0; c < k; c++)
				sub[r * k + c] = g < k ? (uint8_t)(g == c) : P[(g - k) * k + c];
		}
		if (rs_oracle_invert(sub, k)) { /* cannot happen for an MDS matrix */
			if (status) status[s] = -1;
			bad++;
			continue;
		}
		/* step 1: missing data shards from the k survivors */
		const uint8_t *in[256];
		uint8_t *out[256];
		int nrows = 0;
		for (int r = 0; r < k; r++) in[r] = base + (size_t)surv[r] * stride;
		for (int d = 0; d < k; d++)
			if (!pr[d]) {
				memcpy(rows + nrows * k, sub + d * k, k);
				out[nrows++] = base + (size_t)d * stride;
			}
		if (nrows) encode_rows(k, nrows, rows, in, out, len);
		/* step 2: missing parity shards re-encoded from the (now complete) data */
		nrows = 0;
		for (int d = 0; d < k; d++) in[d] = base + (size_t)d * stride;
		for (int i = 0; i < m; i++)
			if (!pr[k + i]) {
				memcpy(rows + nrows * k, P + i * k, k);
				out[nrows++] = base + (size_t)(k + i) * stride;
			}
		if (nrows) encode_rows(k, nrows, rows, in, out, len);
	}
	free(sub);
	free(rows);
	return bad;
}

void rs_oracle_verify(int k, int m, const uint8_t *P, const uint8_t *shards, uint32_t *mismatch,
		      const uint32_t *shard_len, size_t stride, size_t n)
{
	init_tables();
	const int tot = k + m;
	for (size_t s = 0; s < n; s++) {
		size_t len = shard_len ? shard_len[s] : stride;
		const uint8_t *base = shards + s * tot * stride;
		uint32_t mm = 0;
		for (int i = 0; i < m; i++) {
			const uint8_t *stored = base + (size_t)(k + i) * stride;
			for (size_t t = 0; t < len; t++) {
				uint8_t acc = 0;
				for (int j = 0; j < k; j++)
					acc ^= rs_oracle_gf_mul(P[i * k + j], base[(size_t)j * stride + t]);
				if (acc != stored[t]) { mm |= 1u << i; break; }
			}
		}
		mismatch[s] = mm;
	}
}

uint32_t rs_oracle_shard_len(uint32_t block_len, int k)
{
	return (block_len + (uint32_t)k - 1) / (uint32_t)k;
}

void rs_oracle_split_block(const uint8_t *block, uint32_t block_len, int k, uint8_t *dst,
			   size_t stride)
{
	uint32_t L = rs_oracle_shard_len(block_len, k);
	for (int j = 0; j < k; j++) {
		size_t off = (size_t)j * L;
		size_t have = off < block_len ? block_len - off : 0;
		if (have > L) have = L;
		memcpy(dst + (size_t)j * stride, block + off, have);
		memset(dst + (size_t)j * stride + have, 0, L - have);
	}
}

void rs_oracle_join_block(const uint8_t *shards, uint32_t block_len, int k, size_t stride,
			  uint8_t *block)
{
	uint32_t L = rs_oracle_shard_len(block_len, k);
	for (int j = 0; j < k; j++) {
		size_t off = (size_t)j * L;
		size_t have = off < block_len ? block_len - off : 0;
		if (have > L) have = L;
		memcpy(block + off, shards + (size_t)j * stride, have);
	}
}

static inline uint64_t splitmix64_at(uint64_t seed, uint64_t idx)
{
	uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

void rs_oracle_fill_random(uint8_t *dst, size_t len, uint64_t seed, uint64_t offset)
{
	uint64_t idx = offset / 8;
	size_t t = 0;
	while (t < len) {
		uint64_t w = splitmix64_at(seed, idx++);
		for (int b = 0; b < 8 && t < len; b++, t++) dst[t] = (uint8_t)(w >> (8 * b));
	}
}

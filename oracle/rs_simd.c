/*
 * rs_simd.c -- the "fast CPU" restatement of the RS block path: AVX-512 + GFNI
 * (vgf2p8affineqb), AVX2 PSHUFB split-nibble, or a scalar product table, chosen at run
 * time from cpuid, with stripes spread over a pthread pool.
 *
 * TEST INFRASTRUCTURE ONLY (see rs_oracle.h): this is the CPU arm that bench.py times
 * beside the GPU (cpu_baseline / --impl reference).  PARITY UNPINNED by the reference (it
 * ships no Reed-Solomon, doc/book/design/goals.md:27); tests/test_oracle.py requires this
 * file to agree byte for byte with the scalar oracle rs_oracle.c on every ISA path.
 *
 * It stands in for what BASELINE.json calls the reference's "Rust/SIMD Reed-Solomon":
 * the same technique class (constant-multiply by SIMD table/affine, all host cores) that
 * the reed-solomon-erasure crate's simd-accel feature / ISA-L use.
 */
#define _GNU_SOURCE
#include "rs_oracle.h"

#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

enum { ISA_SCALAR = 0, ISA_AVX2 = 1, ISA_GFNI512 = 2 };
static int g_isa = -1;
static int g_force_isa = -1;

static int detect_isa(void)
{
	if (g_force_isa >= 0) return g_force_isa;
	__builtin_cpu_init();
	if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
	    __builtin_cpu_supports("gfni"))
		return ISA_GFNI512;
	if (__builtin_cpu_supports("avx2")) return ISA_AVX2;
	return ISA_SCALAR;
}

static int isa(void)
{
	if (g_isa < 0) g_isa = detect_isa();
	return g_isa;
}

/* test hook: force an ISA path (-1 = auto). returns the path now in use; a request for
 * an ISA the CPU lacks is refused (stays on auto). */
int rs_simd_force_isa(int want)
{
	int best;
	g_force_isa = -1;
	best = detect_isa();
	if (want >= 0 && want <= best) g_force_isa = want;
	g_isa = detect_isa();
	return g_isa;
}

const char *rs_simd_isa(void)
{
	switch (isa()) {
	case ISA_GFNI512: return "avx512+gfni";
	case ISA_AVX2: return "avx2-pshufb";
	default: return "scalar-table";
	}
}

int rs_simd_max_threads(void)
{
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof(set), &set) == 0) {
		int c = CPU_COUNT(&set);
		if (c > 0) return c;
	}
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	return n > 0 ? (int)n : 1;
}

/* ------------------------------------------------------------------ scalar path */
static void rows_scalar(int k, int rows, const uint8_t *C, const uint8_t *const *in,
			uint8_t *const *out, size_t len)
{
	uint8_t tab[256];
	for (int i = 0; i < rows; i++) {
		uint8_t *o = out[i];
		for (int j = 0; j < k; j++) {
			uint8_t c = C[i * k + j];
			const uint8_t *d = in[j];
			for (int x = 0; x < 256; x++) tab[x] = rs_oracle_gf_mul(c, (uint8_t)x);
			if (j == 0)
				for (size_t t = 0; t < len; t++) o[t] = tab[d[t]];
			else
				for (size_t t = 0; t < len; t++) o[t] ^= tab[d[t]];
		}
	}
}

/* ------------------------------------------------------------------ AVX2 path */
/* prepared coefficient tables for one (rows x k) matrix */
struct prep {
	int k, rows;
	const uint8_t *C;
	uint8_t (*lo)[16], (*hi)[16]; /* AVX2: split-nibble tables c*x = lo[x&15] ^ hi[x>>4] */
	uint64_t *A;                   /* GFNI: 8x8 bit matrices */
};

static uint64_t affine_matrix_for(uint8_t c);

static void prep_init(struct prep *p, int k, int rows, const uint8_t *C)
{
	memset(p, 0, sizeof(*p));
	p->k = k;
	p->rows = rows;
	p->C = C;
	if (isa() == ISA_AVX2) {
		p->lo = malloc((size_t)rows * k * 16);
		p->hi = malloc((size_t)rows * k * 16);
		for (int e = 0; e < rows * k; e++)
			for (int x = 0; x < 16; x++) {
				p->lo[e][x] = rs_oracle_gf_mul(C[e], (uint8_t)x);
				p->hi[e][x] = rs_oracle_gf_mul(C[e], (uint8_t)(x << 4));
			}
	} else if (isa() == ISA_GFNI512) {
		p->A = malloc((size_t)rows * k * sizeof(uint64_t));
		for (int e = 0; e < rows * k; e++) p->A[e] = affine_matrix_for(C[e]);
	}
}

static void prep_free(struct prep *p)
{
	free(p->lo);
	free(p->hi);
	free(p->A);
	memset(p, 0, sizeof(*p));
}

__attribute__((target("avx2"))) static void rows_avx2(const struct prep *p,
							const uint8_t *const *in,
							uint8_t *const *out, size_t len)
{
	const int k = p->k, rows = p->rows;
	const uint8_t *C = p->C;
	uint8_t(*lo)[16] = p->lo;
	uint8_t(*hi)[16] = p->hi;
	const __m256i mask = _mm256_set1_epi8(0x0f);
	size_t t = 0;
	for (; t + 32 <= len; t += 32) {
		for (int i0 = 0; i0 < rows; i0 += 4) {
			int nr = rows - i0 < 4 ? rows - i0 : 4;
			__m256i acc[4];
			for (int r = 0; r < nr; r++) acc[r] = _mm256_setzero_si256();
			for (int j = 0; j < k; j++) {
				__m256i d = _mm256_loadu_si256((const __m256i *)(in[j] + t));
				__m256i dl = _mm256_and_si256(d, mask);
				__m256i dh = _mm256_and_si256(_mm256_srli_epi64(d, 4), mask);
				for (int r = 0; r < nr; r++) {
					int e = (i0 + r) * k + j;
					__m256i tl = _mm256_broadcastsi128_si256(
						_mm_loadu_si128((const __m128i *)lo[e]));
					__m256i th = _mm256_broadcastsi128_si256(
						_mm_loadu_si128((const __m128i *)hi[e]));
					acc[r] = _mm256_xor_si256(
						acc[r], _mm256_xor_si256(_mm256_shuffle_epi8(tl, dl),
									 _mm256_shuffle_epi8(th, dh)));
				}
			}
			for (int r = 0; r < nr; r++)
				_mm256_storeu_si256((__m256i *)(out[i0 + r] + t), acc[r]);
		}
	}
	if (t < len) {
		const uint8_t *in2[256];
		uint8_t *out2[256];
		for (int j = 0; j < k; j++) in2[j] = in[j] + t;
		for (int i = 0; i < rows; i++) out2[i] = out[i] + t;
		rows_scalar(k, rows, C, in2, out2, len - t);
	}
}

/* ------------------------------------------------------------------ AVX-512 + GFNI path */
static uint64_t affine_matrix_for(uint8_t c)
{
	/* vgf2p8affineqb: out bit i = parity(matrix.byte[7-i] & x).  For y = c*x,
	 * y bit i = XOR over input bits b with bit i of (c * 2^b) set. */
	uint64_t mat = 0;
	for (int i = 0; i < 8; i++) {
		uint8_t row = 0;
		for (int b = 0; b < 8; b++)
			if ((rs_oracle_gf_mul(c, (uint8_t)(1u << b)) >> i) & 1) row |= (uint8_t)(1u << b);
		mat |= (uint64_t)row << (8 * (7 - i));
	}
	return mat;
}

__attribute__((target("avx512f,avx512bw,gfni"))) static void
rows_gfni512(const struct prep *p, const uint8_t *const *in, uint8_t *const *out, size_t len)
{
	const int k = p->k, rows = p->rows;
	const uint64_t *A = p->A;
	for (size_t t = 0; t < len; t += 64) {
		size_t rem = len - t;
		__mmask64 msk = rem >= 64 ? ~(__mmask64)0 : (((__mmask64)1 << rem) - 1);
		for (int i0 = 0; i0 < rows; i0 += 4) {
			int nr = rows - i0 < 4 ? rows - i0 : 4;
			__m512i acc[4];
			for (int r = 0; r < nr; r++) acc[r] = _mm512_setzero_si512();
			for (int j = 0; j < k; j++) {
				__m512i d = _mm512_maskz_loadu_epi8(msk, in[j] + t);
				for (int r = 0; r < nr; r++) {
					__m512i a = _mm512_set1_epi64((long long)A[(i0 + r) * k + j]);
					acc[r] = _mm512_xor_si512(
						acc[r], _mm512_gf2p8affine_epi64_epi8(d, a, 0));
				}
			}
			for (int r = 0; r < nr; r++)
				_mm512_mask_storeu_epi8(out[i0 + r] + t, msk, acc[r]);
		}
	}
}

/* out[i] = XOR_j C[i][j] * in[j]  over len bytes, best ISA */
static void rows_apply(const struct prep *p, const uint8_t *const *in, uint8_t *const *out,
		       size_t len)
{
	switch (isa()) {
	case ISA_GFNI512: rows_gfni512(p, in, out, len); break;
	case ISA_AVX2: rows_avx2(p, in, out, len); break;
	default: rows_scalar(p->k, p->rows, p->C, in, out, len); break;
	}
}

void rs_simd_rows(int k, int rows, const uint8_t *C, const uint8_t *const *in, uint8_t *const *out,
		  size_t len)
{
	struct prep p;
	prep_init(&p, k, rows, C);
	rows_apply(&p, in, out, len);
	prep_free(&p);
}

/* ------------------------------------------------------------------ threaded batch drivers */
struct job {
	int op; /* 0 encode, 1 reconstruct, 2 verify, 3 copy (first touch) */
	const uint8_t *cp_src;
	uint8_t *cp_dst;
	size_t cp_bytes; /* per stripe */
	int k, m;
	const uint8_t *P;
	const uint8_t *data;
	uint8_t *parity;
	uint8_t *shards;
	const uint8_t *present;
	int32_t *status;
	uint32_t *mismatch;
	const uint32_t *shard_len;
	size_t stride, s0, s1;
	size_t bad;
};

/* segment of columns processed per pass so the k inputs + m outputs stay in L1/L2 */
#define SEG 4096

static void apply_seg(const struct prep *p, const uint8_t *const *in, uint8_t *const *out,
		      size_t len)
{
	const uint8_t *in2[256];
	uint8_t *out2[256];
	for (size_t t = 0; t < len; t += SEG) {
		size_t n = len - t < SEG ? len - t : SEG;
		for (int j = 0; j < p->k; j++) in2[j] = in[j] + t;
		for (int i = 0; i < p->rows; i++) out2[i] = out[i] + t;
		rows_apply(p, in2, out2, n);
	}
}

static void encode_one(int k, int rows, const uint8_t *C, const uint8_t *const *in,
		       uint8_t *const *out, size_t len)
{
	struct prep p;
	prep_init(&p, k, rows, C);
	apply_seg(&p, in, out, len);
	prep_free(&p);
}

static void *worker(void *arg)
{
	struct job *jb = (struct job *)arg;
	if (jb->op == 3) {
		for (size_t s = jb->s0; s < jb->s1; s++)
			memcpy(jb->cp_dst + s * jb->cp_bytes, jb->cp_src + s * jb->cp_bytes, jb->cp_bytes);
		return NULL;
	}
	const int k = jb->k, m = jb->m, tot = k + m;
	const uint8_t *in[256];
	uint8_t *out[256];
	uint8_t *sub = malloc((size_t)k * k), *rows = malloc((size_t)tot * k);
	uint8_t *tmp = NULL;
	struct prep pP;
	prep_init(&pP, k, m, jb->P);
	if (jb->op == 2) tmp = malloc((size_t)m * SEG);
	for (size_t s = jb->s0; s < jb->s1; s++) {
		size_t len = jb->shard_len ? jb->shard_len[s] : jb->stride;
		if (jb->op == 0) {
			for (int j = 0; j < k; j++) in[j] = jb->data + (s * k + j) * jb->stride;
			for (int i = 0; i < m; i++) out[i] = jb->parity + (s * m + i) * jb->stride;
			apply_seg(&pP, in, out, len);
		} else if (jb->op == 1) {
			const uint8_t *pr = jb->present + s * tot;
			uint8_t *base = jb->shards + s * tot * jb->stride;
			int surv[256], ns = 0, np = 0, nrows = 0;
			for (int i = 0; i < tot; i++) np += pr[i] ? 1 : 0;
			if (np < k) {
				if (jb->status) jb->status[s] = -1;
				jb->bad++;
				continue;
			}
			if (jb->status) jb->status[s] = 0;
			if (np == tot) continue;
			for (int i = 0; i < tot && ns < k; i++)
				if (pr[i]) surv[ns++] = i;
			for (int r = 0; r < k; r++)
				for (int c = 0; c < k; c++)
					sub[r * k + c] = surv[r] < k ? (uint8_t)(surv[r] == c)
								     : jb->P[(surv[r] - k) * k + c];
			if (rs_oracle_invert(sub, k)) {
				if (jb->status) jb->status[s] = -1;
				jb->bad++;
				continue;
			}
			for (int r = 0; r < k; r++) in[r] = base + (size_t)surv[r] * jb->stride;
			for (int d = 0; d < k; d++)
				if (!pr[d]) {
					memcpy(rows + nrows * k, sub + d * k, k);
					out[nrows++] = base + (size_t)d * jb->stride;
				}
			if (nrows) encode_one(k, nrows, rows, in, out, len);
			nrows = 0;
			for (int d = 0; d < k; d++) in[d] = base + (size_t)d * jb->stride;
			for (int i = 0; i < m; i++)
				if (!pr[k + i]) {
					memcpy(rows + nrows * k, jb->P + i * k, k);
					out[nrows++] = base + (size_t)(k + i) * jb->stride;
				}
			if (nrows) encode_one(k, nrows, rows, in, out, len);
		} else {
			const uint8_t *base = jb->shards + s * tot * jb->stride;
			uint32_t mm = 0;
			for (size_t t = 0; t < len; t += SEG) {
				size_t n = len - t < SEG ? len - t : SEG;
				for (int j = 0; j < k; j++) in[j] = base + (size_t)j * jb->stride + t;
				for (int i = 0; i < m; i++) out[i] = tmp + (size_t)i * SEG;
				rows_apply(&pP, in, out, n);
				for (int i = 0; i < m; i++)
					if (memcmp(out[i], base + (size_t)(k + i) * jb->stride + t, n))
						mm |= 1u << i;
			}
			jb->mismatch[s] = mm;
		}
	}
	prep_free(&pP);
	free(sub);
	free(rows);
	free(tmp);
	return NULL;
}

/* Persistent worker pool: creating 128 pthreads per call costs milliseconds, which is a large part
 * of a call on a big host and would understate the CPU arm.  Workers sleep on a condition variable;
 * a call publishes `threads` job slots, workers (and the caller) grab slots until none are left. */
static struct {
	pthread_mutex_t mu;
	pthread_cond_t work, done;
	int nthreads;          /* workers created so far */
	struct job *jobs;      /* current batch */
	int njobs, next, remaining;
	unsigned long epoch;
} g_pool = { PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, 0, NULL, 0, 0, 0, 0 };
static pthread_mutex_t g_call_mu = PTHREAD_MUTEX_INITIALIZER; /* one batch at a time */

static void *pool_main(void *arg)
{
	(void)arg;
	pthread_mutex_lock(&g_pool.mu);
	for (;;) {
		while (g_pool.next >= g_pool.njobs) pthread_cond_wait(&g_pool.work, &g_pool.mu);
		struct job *jb = &g_pool.jobs[g_pool.next++];
		pthread_mutex_unlock(&g_pool.mu);
		worker(jb);
		pthread_mutex_lock(&g_pool.mu);
		if (--g_pool.remaining == 0) pthread_cond_broadcast(&g_pool.done);
	}
	return NULL;
}

static size_t run_jobs(struct job *proto, size_t n, int threads)
{
	if (threads < 1) threads = rs_simd_max_threads();
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	rs_oracle_gf_mul(1, 1); /* make sure the oracle tables are built before fan-out */
	(void)isa();
	struct job *jobs = calloc((size_t)threads, sizeof(*jobs));
	size_t bad = 0;
	for (int t = 0; t < threads; t++) {
		jobs[t] = *proto;
		jobs[t].s0 = n * (size_t)t / (size_t)threads;
		jobs[t].s1 = n * (size_t)(t + 1) / (size_t)threads;
		jobs[t].bad = 0;
	}
	pthread_mutex_lock(&g_call_mu);
	pthread_mutex_lock(&g_pool.mu);
	while (g_pool.nthreads < threads - 1) { /* the caller is the last worker */
		pthread_t th;
		pthread_attr_t at;
		pthread_attr_init(&at);
		pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
		if (pthread_create(&th, &at, pool_main, NULL) != 0) {
			pthread_attr_destroy(&at);
			break;
		}
		pthread_attr_destroy(&at);
		g_pool.nthreads++;
	}
	g_pool.jobs = jobs;
	g_pool.njobs = threads;
	g_pool.next = 0;
	g_pool.remaining = threads;
	pthread_cond_broadcast(&g_pool.work);
	/* the caller works too */
	while (g_pool.next < g_pool.njobs) {
		struct job *jb = &g_pool.jobs[g_pool.next++];
		pthread_mutex_unlock(&g_pool.mu);
		worker(jb);
		pthread_mutex_lock(&g_pool.mu);
		if (--g_pool.remaining == 0) pthread_cond_broadcast(&g_pool.done);
	}
	while (g_pool.remaining > 0) pthread_cond_wait(&g_pool.done, &g_pool.mu);
	g_pool.njobs = 0;
	g_pool.next = 0;
	g_pool.jobs = NULL;
	pthread_mutex_unlock(&g_pool.mu);
	pthread_mutex_unlock(&g_call_mu);
	for (int t = 0; t < threads; t++) bad += jobs[t].bad;
	free(jobs);
	return bad;
}

/* same geometry and semantics as rs_oracle_encode / _reconstruct / _verify; threads<=0 = all */
void rs_simd_encode(int k, int m, const uint8_t *P, const uint8_t *data, uint8_t *parity,
		    const uint32_t *shard_len, size_t stride, size_t n, int threads)
{
	struct job jb = { .op = 0, .k = k, .m = m, .P = P, .data = data, .parity = parity,
			  .shard_len = shard_len, .stride = stride };
	run_jobs(&jb, n, threads);
}

size_t rs_simd_reconstruct(int k, int m, const uint8_t *P, uint8_t *shards, const uint8_t *present,
			   int32_t *status, const uint32_t *shard_len, size_t stride, size_t n,
			   int threads)
{
	struct job jb = { .op = 1, .k = k, .m = m, .P = P, .shards = shards, .present = present,
			  .status = status, .shard_len = shard_len, .stride = stride };
	return run_jobs(&jb, n, threads);
}

void rs_simd_verify(int k, int m, const uint8_t *P, const uint8_t *shards, uint32_t *mismatch,
		    const uint32_t *shard_len, size_t stride, size_t n, int threads)
{
	struct job jb = { .op = 2, .k = k, .m = m, .P = P, .shards = (uint8_t *)shards,
			  .mismatch = mismatch, .shard_len = shard_len, .stride = stride };
	run_jobs(&jb, n, threads);
}

/* NUMA first touch: copies n stripes of `bytes_per_stripe` with the same stripe->thread partition
 * the drivers above use, so that a fresh (untouched) destination ends up on the memory node of the
 * thread that will process it.  Without it a 128-thread host streams everything from the node of
 * the thread that generated the data (measured: 12 GiB/s instead of > 40). */
void rs_simd_parallel_copy(uint8_t *dst, const uint8_t *src, size_t bytes_per_stripe, size_t n, int threads)
{
	struct job jb = { .op = 3, .cp_src = src, .cp_dst = dst, .cp_bytes = bytes_per_stripe };
	run_jobs(&jb, n, threads);
}

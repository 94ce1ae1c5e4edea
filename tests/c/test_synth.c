/* tests/c/test_synth.c -- kmer_cuda_dev_synth_reads (csrc/synth.cu) from a plain C host against a C restatement of the
 * generator: base g of the table = "ACGT"[(splitmix64(seed + (g/32 + 1) * 0x9E3779B97F4A7C15) >> 2*(g%32)) & 3]; then the
 * generated column is counted by kmer_cuda_dev_count WITHOUT leaving the device and the table is checked against the C
 * oracle's GROUP BY of the restated rows.  Built and run by tests/test_synth_gpu.py.  Exit code 0 = bit-exact. */
#include <cuda_runtime_api.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmer_cuda.h"

int orc_count(const char *flat, const uint64_t *off, uint64_t n_rows, int k, uint64_t *keys, uint64_t *counts,
			  uint64_t *n_distinct, uint64_t *n_kmers, int64_t *bad_row);

static uint64_t splitmix(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static char base_at(uint64_t seed, uint64_t g)
{
	return "ACGT"[(splitmix(seed + ((g >> 5) + 1) * 0x9E3779B97F4A7C15ULL) >> (2 * (g & 31))) & 3];
}

static int cmp_pair(const void *a, const void *b)
{
	uint64_t x = ((const kmer_count_pair *) a)->code, y = ((const kmer_count_pair *) b)->code;
	return x < y ? -1 : x > y;
}

int main(void)
{
	kmer_cuda_ctx *ctx = NULL;
	if (kmer_cuda_init(&ctx, 0)) { printf("init failed: %s\n", kmer_cuda_last_error(NULL)->message); return 2; }
	int failures = 0;
	static const struct { uint64_t seed, first, rows, len; int k; } cases[] = {
		{2, 0, 3000, 1000, 21}, {3, 12345, 2000, 997, 31}, {0xFFFFFFFFFFFFFFFFULL, 7, 301, 75, 5}, {5, 1, 1, 15, 14}, {6, 3, 0, 100, 21},
		{7, 1000003, 40, 33, 32}};
	for (unsigned ci = 0; ci < sizeof(cases) / sizeof(cases[0]); ci++)
	{
		uint64_t seed = cases[ci].seed, first = cases[ci].first, rows = cases[ci].rows, len = cases[ci].len, n = rows * len;
		int k = cases[ci].k;
		char *d_seq = NULL;
		uint64_t *d_off = NULL;
		kmer_count_pair *d_pairs = NULL;
		uint64_t cap = kmer_cuda_max_kmers(n, rows, k) + 1;
		if (cudaMalloc((void **) &d_seq, ((n + 15) & ~15ULL) + 64) || cudaMalloc((void **) &d_off, (rows + 1) * 8) ||
			cudaMalloc((void **) &d_pairs, cap * sizeof(kmer_count_pair))) { printf("cudaMalloc failed\n"); return 2; }
		cudaMemset(d_seq, 0, ((n + 15) & ~15ULL) + 64);
		kmer_dev_result res;
		int rc = kmer_cuda_dev_synth_reads(ctx, seed, first, rows, len, d_seq, d_off, NULL);
		if (!rc) rc = kmer_cuda_dev_count(ctx, d_seq, n, d_off, rows, k, d_pairs, cap, 0, NULL);
		if (!rc) rc = kmer_cuda_dev_finish(ctx, NULL, &res);
		if (rc) { printf("case %u: failed (%d): %s\n", ci, rc, kmer_cuda_last_error(ctx)->message); failures++; continue; }
		char *got = malloc(n + 1), *want = malloc(n + 1);
		uint64_t *off = malloc((rows + 1) * 8), *woff = malloc((rows + 1) * 8);
		cudaMemcpy(got, d_seq, n, cudaMemcpyDeviceToHost);
		cudaMemcpy(off, d_off, (rows + 1) * 8, cudaMemcpyDeviceToHost);
		for (uint64_t i = 0; i < n; i++) want[i] = base_at(seed, first * len + i);
		for (uint64_t r = 0; r <= rows; r++) woff[r] = r * len;
		int ok = !memcmp(got, want, n) && !memcmp(off, woff, (rows + 1) * 8);
		/* the generated column counted in place == the oracle's table of the restated rows */
		uint64_t *wk = malloc((n + 1) * 8), *wc = malloc((n + 1) * 8), wd = 0, wn = 0;
		int64_t bad;
		if (orc_count(want, woff, rows, k, wk, wc, &wd, &wn, &bad)) { printf("oracle error\n"); return 2; }
		kmer_count_pair *tab = malloc((res.n_distinct + 1) * sizeof(kmer_count_pair));
		cudaMemcpy(tab, d_pairs, res.n_distinct * sizeof(kmer_count_pair), cudaMemcpyDeviceToHost);
		qsort(tab, res.n_distinct, sizeof(kmer_count_pair), cmp_pair);
		int okc = res.n_distinct == wd && res.n_kmers == wn;
		for (uint64_t i = 0; okc && i < wd; i++) okc = tab[i].code == wk[i] && tab[i].count == wc[i];
		printf("[synth] seed=%llu first_row=%llu %llu x %llu k=%d: text %s, count %s (groups %llu/%llu)\n", (unsigned long long) seed,
			   (unsigned long long) first, (unsigned long long) rows, (unsigned long long) len, k, ok ? "ok" : "MISMATCH", okc ? "ok" : "MISMATCH",
			   (unsigned long long) res.n_distinct, (unsigned long long) wd);
		if (!ok || !okc) failures++;
		free(got); free(want); free(off); free(woff); free(wk); free(wc); free(tab);
		cudaFree(d_seq); cudaFree(d_off); cudaFree(d_pairs);
	}
	kmer_cuda_shutdown(ctx);
	return failures ? 1 : 0;
}

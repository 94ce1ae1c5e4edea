/* tests/c/test_multi.c -- the multi-GPU C ABI from a plain C host (no Python, no torch, no NCCL):
 * kmer_cuda_init_multi + kmer_cuda_multi_submit_count on the devices given on the command line ("0,1"; "0,0" runs two
 * contexts on one GPU) against the C oracle's GROUP BY table (oracle/kmer_oracle.c orc_count).  Built and run by
 * tests/test_multi_c_abi.py.  Exit code 0 = every case bit-exact. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmer_cuda.h"

int orc_count(const char *flat, const uint64_t *off, uint64_t n_rows, int k, uint64_t *keys, uint64_t *counts,
			  uint64_t *n_distinct, uint64_t *n_kmers, int64_t *bad_row);

static uint64_t rng_state = 88172645463325252ULL;
static uint32_t rnd(void)
{
	rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
	return (uint32_t) (rng_state >> 32);
}

static int cmp_pair(const void *a, const void *b)
{
	uint64_t x = ((const kmer_count_pair *) a)->code, y = ((const kmer_count_pair *) b)->code;
	return x < y ? -1 : x > y;
}

/* skew: 0 = random reads of ragged length; 1 = poly-A / poly-T rows and one read repeated many times mixed in */
static void make_rows(int n_rows, int skew, char **flat_out, uint64_t **off_out)
{
	uint64_t cap = (uint64_t) n_rows * 900 + 16, n = 0;
	char *flat = malloc(cap);
	uint64_t *off = malloc(((size_t) n_rows + 1) * 8);
	char rep[400];
	for (int i = 0; i < 400; i++) rep[i] = "ACGT"[rnd() & 3];
	off[0] = 0;
	for (int r = 0; r < n_rows; r++)
	{
		int kind = skew ? r % 5 : 4;
		int len = 64 + (int) (rnd() % 700);
		if (kind == 0) { for (int i = 0; i < len; i++) flat[n++] = 'A'; }
		else if (kind == 1) { memcpy(flat + n, rep, 400); n += 400; }
		else if (kind == 2) { for (int i = 0; i < 64; i++) flat[n++] = 't'; for (int i = 0; i < len; i++) flat[n++] = "ACGT"[rnd() & 3]; }
		else { for (int i = 0; i < len; i++) flat[n++] = "ACGTacgt"[rnd() & 7]; }
		off[r + 1] = n;
	}
	*flat_out = flat;
	*off_out = off;
}

int main(int argc, char **argv)
{
	int devices[16], nd = 0, failures = 0;
	const char *list = argc > 1 ? argv[1] : "0,0";
	for (const char *p = list; *p && nd < 16;)
	{
		devices[nd++] = atoi(p);
		p = strchr(p, ',');
		if (!p) break;
		p++;
	}
	kmer_cuda_multi *m = NULL;
	int rc = kmer_cuda_init_multi(&m, devices, nd);
	if (rc)
	{
		printf("init_multi failed: %s\n", kmer_cuda_multi_last_error(NULL)->message);
		return 2;
	}
	static const int ks[] = {21, 31, 14, 32, 5, 13};
	for (int skew = 0; skew < 2; skew++)
	{
		char *flat;
		uint64_t *off;
		int n_rows = skew ? 3000 : 6000;
		make_rows(n_rows, skew, &flat, &off);
		for (unsigned ki = 0; ki < sizeof(ks) / sizeof(ks[0]); ki++)
		{
			int k = ks[ki];
			uint64_t cap = off[n_rows], want_d = 0, want_n = 0, got_n = 0, got_d = 0;
			int64_t bad;
			uint64_t *wk = malloc(cap * 8), *wc = malloc(cap * 8);
			if (orc_count(flat, off, n_rows, k, wk, wc, &want_d, &want_n, &bad)) { printf("oracle error\n"); return 2; }
			kmer_count_pair *pairs[16];
			uint64_t n_distinct[16];
			rc = kmer_cuda_multi_submit_count(m, flat, off, n_rows, k, pairs, n_distinct, &got_n);
			if (rc)
			{
				printf("k=%d skew=%d: submit failed (%d): %s\n", k, skew, rc, kmer_cuda_multi_last_error(m)->message);
				failures++;
				continue;
			}
			for (int d = 0; d < nd; d++) got_d += n_distinct[d];
			kmer_count_pair *all = malloc((got_d ? got_d : 1) * sizeof(kmer_count_pair));
			uint64_t o = 0;
			for (int d = 0; d < nd; d++)
			{
				memcpy(all + o, pairs[d], n_distinct[d] * sizeof(kmer_count_pair));
				o += n_distinct[d];
				kmer_cuda_multi_release(m, d, pairs[d]);
			}
			qsort(all, got_d, sizeof(kmer_count_pair), cmp_pair);
			int ok = got_d == want_d && got_n == want_n;
			for (uint64_t i = 0; ok && i < want_d; i++) ok = all[i].code == wk[i] && all[i].count == wc[i];
			printf("[multi x%d] k=%d %s rows: %s  groups=%llu/%llu k-mers=%llu (per device:", nd, k, skew ? "skewed" : "random", ok ? "ok" : "MISMATCH",
				   (unsigned long long) got_d, (unsigned long long) want_d, (unsigned long long) got_n);
			for (int d = 0; d < nd; d++) printf(" %llu", (unsigned long long) n_distinct[d]);
			printf(")\n");
			if (!ok) failures++;
			free(all); free(wk); free(wc);
		}
		free(flat); free(off);
	}
	/* errors: the first offending row of the whole batch, as a sequential scan would report it */
	{
		const char *rows = "ACGTACGTACGTACGTACGTACGTAAAAACGTACGTACGTACGTACGTACGTAAAAACGTNCGTACGTACGTACGTACGTAAAAACGT";
		uint64_t off[4] = {0, 29, 58, 87};
		kmer_count_pair *pairs[16];
		uint64_t n_distinct[16], nk;
		rc = kmer_cuda_multi_submit_count(m, rows, off, 3, 21, pairs, n_distinct, &nk);
		const kmer_cuda_error *e = kmer_cuda_multi_last_error(m);
		int ok = rc == KMER_ERR_INVALID_DNA && e->row == 2 && !strcmp(e->sqlstate, "22P02") && !strcmp(e->message, "Invalid DNA Sequence");
		printf("[multi x%d] invalid character in row 2: %s (rc=%d row=%lld %s)\n", nd, ok ? "ok" : "MISMATCH", rc, (long long) e->row, e->message);
		if (!ok) failures++;
		rc = kmer_cuda_multi_submit_count(m, rows, off, 3, 30, pairs, n_distinct, &nk);
		e = kmer_cuda_multi_last_error(m);
		ok = rc == KMER_ERR_INVALID_K && e->row == 0 && !strcmp(e->sqlstate, "22023");
		printf("[multi x%d] rows shorter than k: %s (rc=%d row=%lld %s)\n", nd, ok ? "ok" : "MISMATCH", rc, (long long) e->row, e->message);
		if (!ok) failures++;
	}
	kmer_cuda_shutdown_multi(m);
	return failures ? 1 : 0;
}

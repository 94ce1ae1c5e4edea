/* tests/c/test_multi_match.c -- kmer_cuda_multi_submit_match from a plain C host (no Python, no torch, no NCCL): a k-mer
 * column sharded over the devices given on the command line ("0,1"; "0,0" = two contexts on one GPU; "0" = one device),
 * constants replicated, against the C oracle's per-pair predicates (oracle/kmer_oracle.c orc_match_column: kmer_equals
 * kmer.c:226-245, kmer_starts_with kmer.c:248-265, kmer_contains kmer.c:268-285).  Built and run by
 * tests/test_sharded_match_gpu.py.  Exit code 0 = every bit and every hit count exact. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmer_cuda.h"

int orc_match_column(int op, const uint64_t *codes, const uint8_t *lens, uint64_t m, int k, const char *text, uint8_t *out);

static uint64_t rng_state = 0x9E3779B97F4A7C15ULL;
static uint64_t rnd64(void)
{
	rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
	return rng_state;
}

static void code_text(uint64_t code, int k, char *out)
{
	for (int j = 0; j < k; j++) out[j] = "acgt"[(code >> (2 * (k - 1 - j))) & 3];
	out[k] = 0;
}

/* one case: m k-mers of length k (or mixed lengths when lens != NULL) against n_c constants with per-constant ops */
static int run_case(kmer_cuda_multi *mh, int nd, const char *name, const uint64_t *codes, const uint8_t *lens, uint64_t m, int k,
					const char **consts, const int *ops, uint32_t n_c)
{
	uint32_t *bits = NULL;
	uint64_t wpr = 0, *hits = NULL;
	int rc = kmer_cuda_multi_submit_match(mh, KMER_OP_EQUALS, ops, codes, lens, m, k, consts, n_c, &bits, &wpr, &hits);
	if (rc)
	{
		printf("[multi-match x%d] %s: submit failed (%d): %s\n", nd, name, rc, kmer_cuda_multi_last_error(mh)->message);
		return 1;
	}
	int ok = wpr == (m + 31) / 32;
	uint8_t *want = malloc(m ? m : 1);
	uint64_t bad_bits = 0;
	for (uint32_t c = 0; ok && c < n_c; c++)
	{
		if (orc_match_column(ops[c], codes, lens, m, k, consts[c], want)) { printf("oracle error\n"); return 1; }
		uint64_t h = 0;
		for (uint64_t i = 0; i < m; i++)
		{
			int got = (bits[c * wpr + (i >> 5)] >> (i & 31)) & 1;
			if (got != want[i]) bad_bits++;
			h += want[i];
		}
		if (bad_bits || hits[c] != h) ok = 0;
	}
	printf("[multi-match x%d] %s: %s  (%llu k-mers x %u constants, wrong bits %llu)\n", nd, name, ok ? "ok" : "MISMATCH",
		   (unsigned long long) m, n_c, (unsigned long long) bad_bits);
	kmer_cuda_multi_release(mh, 0, bits);
	kmer_cuda_multi_release(mh, 0, hits);
	free(want);
	return ok ? 0 : 1;
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
	kmer_cuda_multi *mh = NULL;
	int rc = kmer_cuda_init_multi(&mh, devices, nd);
	if (rc)
	{
		printf("init_multi failed: %s\n", kmer_cuda_multi_last_error(NULL)->message);
		return 2;
	}
	/* 1. k = 12, IUPAC patterns (contains) + an equals and a starts_with constant in the same pass; sizes around the cut rule */
	static const uint64_t sizes[] = {100003, 64, 31, 1, 0, 4096 + 33};
	for (unsigned si = 0; si < sizeof(sizes) / sizeof(sizes[0]); si++)
	{
		uint64_t m = sizes[si];
		int k = 12;
		uint64_t *codes = malloc((m ? m : 1) * 8);
		for (uint64_t i = 0; i < m; i++) codes[i] = rnd64() & ((1ull << (2 * k)) - 1);
		char first[33], prefix[8];
		code_text(m ? codes[0] : 0, k, first);
		memcpy(prefix, first, 5); prefix[5] = 0;
		static char pat[12][13];
		const char *consts[15];
		int ops[15];
		for (int p = 0; p < 12; p++)
		{
			for (int j = 0; j < k; j++) pat[p][j] = "ACGTRYKMSWBDHVNNNNNN"[rnd64() % 20];
			pat[p][k] = 0;
			consts[p] = pat[p];
			ops[p] = KMER_OP_CONTAINS;
		}
		consts[12] = "nnnnnnnnnnnn"; ops[12] = KMER_OP_CONTAINS;
		consts[13] = first; ops[13] = KMER_OP_EQUALS;
		consts[14] = prefix; ops[14] = KMER_OP_STARTS_WITH;
		char name[64];
		snprintf(name, sizeof(name), "k=12 mixed ops, m=%llu", (unsigned long long) m);
		failures += run_case(mh, nd, name, codes, NULL, m, k, consts, ops, 15);
		free(codes);
	}
	/* 2. k = 32 (all 64 bits used): equals + starts_with in one pass */
	{
		uint64_t m = 50021;
		uint64_t *codes = malloc(m * 8);
		for (uint64_t i = 0; i < m; i++) codes[i] = rnd64();
		codes[m - 1] = codes[0];
		char first[33], prefix[9];
		code_text(codes[0], 32, first);
		memcpy(prefix, first, 8); prefix[8] = 0;
		const char *consts[3] = {first, prefix, ""};
		int ops[3] = {KMER_OP_EQUALS, KMER_OP_STARTS_WITH, KMER_OP_STARTS_WITH};   /* the empty prefix matches everything (kmer.c:44-55) */
		failures += run_case(mh, nd, "k=32 equals + starts_with", codes, NULL, m, 32, consts, ops, 3);
		free(codes);
	}
	/* 3. a column of mixed lengths (lens != NULL): length rules of every predicate */
	{
		uint64_t m = 20011;
		uint64_t *codes = malloc(m * 8);
		uint8_t *lens = malloc(m);
		for (uint64_t i = 0; i < m; i++)
		{
			lens[i] = (uint8_t) (rnd64() % 9);            /* 0..8 */
			codes[i] = lens[i] ? rnd64() & ((1ull << (2 * lens[i])) - 1) : 0;
		}
		const char *consts[4] = {"acg", "ac", "nnrya", "t"};
		int ops[4] = {KMER_OP_EQUALS, KMER_OP_STARTS_WITH, KMER_OP_CONTAINS, KMER_OP_EQUALS};
		failures += run_case(mh, nd, "mixed lengths 0..8", codes, lens, m, 0, consts, ops, 4);
		free(codes); free(lens);
	}
	/* 4. a bad constant is reported with its index and the reference's SQLSTATE (qkmer_in kmer.c:149-182) */
	{
		uint64_t codes[40] = {0};
		const char *consts[3] = {"acgt", "acgx", "acgt"};
		uint32_t *bits = NULL;
		uint64_t wpr, *hits = NULL;
		rc = kmer_cuda_multi_submit_match(mh, KMER_OP_CONTAINS, NULL, codes, NULL, 40, 4, consts, 3, &bits, &wpr, &hits);
		const kmer_cuda_error *e = kmer_cuda_multi_last_error(mh);
		int ok = rc == KMER_ERR_INVALID_QKMER && e->row == 1 && !strcmp(e->sqlstate, "22P02") && bits == NULL;
		printf("[multi-match x%d] invalid qkmer constant 1: %s (rc=%d row=%lld %s)\n", nd, ok ? "ok" : "MISMATCH", rc, (long long) e->row, e->message);
		if (!ok) failures++;
	}
	kmer_cuda_shutdown_multi(mh);
	return failures ? 1 : 0;
}

/*
 * oracle/kmer_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference extension's hot path, written from the reference's
 * behaviour (file:line cited per function; reference = NishantSushmakar/kmer-extension).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libkmer_cuda.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the known-answer vectors recorded in the reference's kmer-tests.sql (tests/golden/kat.json)
 *   (2) oracle/_ref/libkmer_ref.so = the reference's own kmer.c compiled unmodified
 *       (live when the .so is present, and through committed fixtures tests/golden/ref_*.json).
 *
 * k-mer <-> code convention shared with include/kmer_cuda.h: a=0 c=1 g=2 t=3, first base in the
 * most significant used bit pair, so that for equal k   code order == memcmp order of the
 * lower-case text the reference stores (kmer.c:28-29, 124-126).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_K 32 /* MAX_KMER_LENGTH, kmer.h:18 */

enum
{
	ORC_OK = 0,
	ORC_INVALID_DNA = 1,	/* kmer.c:31-37  22P02 "Invalid DNA Sequence" */
	ORC_KMER_TOO_LONG = 2,	/* kmer.c:115-120 22001 "KMer Sequence larger than length 32" */
	ORC_INVALID_QKMER = 3,	/* kmer.c:177-182 22P02 "Invalid QKMer Sequence" */
	ORC_INVALID_K = 4,		/* kmer.c:310-313 22023 "Invalid KMER Length" */
	ORC_QKMER_TOO_LONG = 5, /* kmer.c:149-154 22001 "QKMer Sequence larger than length 32" */
};

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; } /* tolower, C locale */

/* validate_sequence, kmer.c:20-41: fold case, accept only a c g t.  Returns -1 if valid, else the
 * index of the first offending byte. */
int64_t orc_validate_dna(const char *s, uint64_t len)
{
	for (uint64_t i = 0; i < len; i++)
	{
		int c = lower((unsigned char) s[i]);
		if (c != 'a' && c != 'c' && c != 'g' && c != 't')
			return (int64_t) i;
	}
	return -1;
}

static int base_code(int c)
{
	switch (lower(c))
	{
	case 'a': return 0;
	case 'c': return 1;
	case 'g': return 2;
	case 't': return 3;
	}
	return -1;
}

/* kmer_in, kmer.c:109-129: length check first (22001), then alphabet (22P02). */
int orc_kmer_encode(const char *s, uint64_t len, uint64_t *code)
{
	if (len > ORC_MAX_K)
		return ORC_KMER_TOO_LONG;
	uint64_t v = 0;
	for (uint64_t i = 0; i < len; i++)
	{
		int b = base_code((unsigned char) s[i]);
		if (b < 0)
			return ORC_INVALID_DNA;
		v = (v << 2) | (uint64_t) b;
	}
	*code = v;
	return ORC_OK;
}

/* kmer_out, kmer.c:131-138: the stored text is lower case. */
void orc_kmer_decode(uint64_t code, int k, char *out)
{
	for (int i = 0; i < k; i++)
		out[i] = "acgt"[(code >> (2 * (k - 1 - i))) & 3];
}

/* qkmer_in, kmer.c:141-190: length check (22001), then the 16-letter alphabet (22P02).
 * out receives the lower-cased pattern. */
int orc_qkmer_parse(const char *s, uint64_t len, char *out)
{
	if (len > ORC_MAX_K)
		return ORC_QKMER_TOO_LONG;
	for (uint64_t i = 0; i < len; i++)
	{
		int c = lower((unsigned char) s[i]);
		if (!strchr("acgturykmswbdhvn", c) || c == 0)
			return ORC_INVALID_QKMER;
		out[i] = (char) c;
	}
	return ORC_OK;
}

/* match(), kmer.h:21-53, restated as set membership: the pattern letter names a set of bases.
 * 'u' is accepted by qkmer_in but names no set in match() (default: false); since a kmer never
 * contains 'u' the `pattern == nucleotide` shortcut cannot fire for it either. */
static int iupac_match(int p, int n)
{
	const char *set;
	switch (p)
	{
	case 'a': set = "a"; break;
	case 'c': set = "c"; break;
	case 'g': set = "g"; break;
	case 't': set = "t"; break;
	case 'n': set = "acgt"; break;
	case 'r': set = "ag"; break;
	case 'y': set = "ct"; break;
	case 'k': set = "gt"; break;
	case 'm': set = "ac"; break;
	case 's': set = "gc"; break;
	case 'w': set = "at"; break;
	case 'b': set = "cgt"; break;
	case 'd': set = "agt"; break;
	case 'h': set = "act"; break;
	case 'v': set = "acg"; break;
	default: set = ""; break;
	}
	return n != 0 && strchr(set, n) != NULL;
}

/* kmer_equals, kmer.c:226-245 */
int orc_equals(uint64_t a, int la, uint64_t b, int lb) { return la == lb && a == b; }

/* kmer_starts_with_helper, kmer.c:44-55: len(prefix) <= len(kmer) and the first len(prefix)
 * characters agree; an empty prefix matches everything. */
int orc_starts_with(uint64_t prefix, int lp, uint64_t kmer, int lk)
{
	if (lp > lk)
		return 0;
	char p[ORC_MAX_K], s[ORC_MAX_K];
	orc_kmer_decode(prefix, lp, p);
	orc_kmer_decode(kmer, lk, s);
	return memcmp(p, s, (size_t) lp) == 0;
}

/* kmer_query, kmer.c:59-79: equal length and every position matches. q is lower-case text. */
int orc_contains(const char *q, int lq, uint64_t kmer, int lk)
{
	if (lq != lk)
		return 0;
	char s[ORC_MAX_K];
	orc_kmer_decode(kmer, lk, s);
	for (int i = 0; i < lq; i++)
		if (!iupac_match(q[i], s[i]))
			return 0;
	return 1;
}

/* Batched forms over a column of k-mers.  lens == NULL means every k-mer has length k.
 * op 0 equals(col, const) ; 1 starts_with(const, col) [= col ^@ const] ; 2 contains(qkmer, col).
 * Returns ORC_OK or the input error of the constant. out[i] in {0,1}. */
int orc_match_column(int op, const uint64_t *codes, const uint8_t *lens, uint64_t m, int k,
					 const char *const_text, uint8_t *out)
{
	uint64_t clen = strlen(const_text), ccode = 0;
	char q[ORC_MAX_K + 1];
	int rc = (op == 2) ? orc_qkmer_parse(const_text, clen, q) : orc_kmer_encode(const_text, clen, &ccode);
	if (rc)
		return rc;
	for (uint64_t i = 0; i < m; i++)
	{
		int lk = lens ? lens[i] : k;
		if (op == 0)
			out[i] = (uint8_t) orc_equals(codes[i], lk, ccode, (int) clen);
		else if (op == 1)
			out[i] = (uint8_t) orc_starts_with(ccode, (int) clen, codes[i], lk);
		else
			out[i] = (uint8_t) orc_contains(q, (int) clen, codes[i], lk);
	}
	return ORC_OK;
}

/* generate_kmers over a table of rows, kmer.c:289-351.  Row r is flat[off[r] .. off[r+1]).
 * Per row: ERROR if len < k or k <= 0 or k > 32 (:310-313); otherwise windows 0 .. len-k in
 * position order (:316, :341-343).  Text reaches generate_kmers through dna_in, so an invalid
 * character is an error too (:85-97).  As in a SQL statement the first error (lowest row) aborts
 * everything.  codes may be NULL to only count.  Returns ORC_* and sets *bad_row. */
int orc_generate(const char *flat, const uint64_t *off, uint64_t n_rows, int k, uint64_t *codes,
				 uint64_t *n_out, int64_t *bad_row)
{
	uint64_t n = 0;
	*bad_row = -1;
	*n_out = 0;
	for (uint64_t r = 0; r < n_rows; r++)
	{
		const char *s = flat + off[r];
		uint64_t len = off[r + 1] - off[r];
		if (orc_validate_dna(s, len) >= 0)
		{
			*bad_row = (int64_t) r;
			return ORC_INVALID_DNA;
		}
		if ((int64_t) len < (int64_t) k || k <= 0 || k > ORC_MAX_K)
		{
			*bad_row = (int64_t) r;
			return ORC_INVALID_K;
		}
		for (uint64_t p = 0; p + (uint64_t) k <= len; p++)
		{
			if (codes)
			{
				uint64_t v = 0;
				for (int j = 0; j < k; j++)
					v = (v << 2) | (uint64_t) base_code((unsigned char) s[p + (uint64_t) j]);
				codes[n] = v;
			}
			n++;
		}
	}
	*n_out = n;
	return ORC_OK;
}

static int cmp_u64(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
	return x < y ? -1 : x > y;
}

/* GROUP BY kmer / count(*) over generate_kmers: exact multiset count keyed by the k-mer
 * (kmer_hash_ops: hash kmer.c:353-365 + equals kmer.c:226-245; aggregation is PostgreSQL core).
 * Result order in PostgreSQL is unspecified; this restatement returns ascending code order.
 * keys/counts must have room for n_kmers entries (upper bound). */
int orc_count(const char *flat, const uint64_t *off, uint64_t n_rows, int k, uint64_t *keys,
			  uint64_t *counts, uint64_t *n_distinct, uint64_t *n_kmers, int64_t *bad_row)
{
	uint64_t n = 0;
	*n_distinct = 0;
	*n_kmers = 0;
	int rc = orc_generate(flat, off, n_rows, k, NULL, &n, bad_row);
	if (rc)
		return rc;
	uint64_t *tmp = (uint64_t *) malloc((n ? n : 1) * sizeof(uint64_t));
	if (!tmp)
		return -1;
	orc_generate(flat, off, n_rows, k, tmp, &n, bad_row);
	qsort(tmp, n, sizeof(uint64_t), cmp_u64);
	uint64_t d = 0;
	for (uint64_t i = 0; i < n;)
	{
		uint64_t j = i;
		while (j < n && tmp[j] == tmp[i])
			j++;
		keys[d] = tmp[i];
		counts[d] = j - i;
		d++;
		i = j;
	}
	free(tmp);
	*n_distinct = d;
	*n_kmers = n;
	return ORC_OK;
}

/* Order-independent checksum of the k-mer multiset generate_kmers (kmer.c:289-351) produces over a table of rows:
 *   *sum = sum over all windows of mix64(code)  (mod 2^64),   *n_out = number of windows.
 * A (k-mer, count) table holds the same multiset iff  sum over groups of count * mix64(code) == *sum  (and its keys are
 * distinct, its counts sum to *n_out).  Used by bench.py / the tests to pin results at sizes where the full table cannot be
 * compared.  mix64 is the murmur3 finaliser (a bijection of 64-bit words). */
static uint64_t orc_mix64(uint64_t x)
{
	x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
	x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
	x ^= x >> 33;
	return x;
}

int orc_multiset_checksum(const char *flat, const uint64_t *off, uint64_t n_rows, int k, uint64_t *sum,
						  uint64_t *n_out, int64_t *bad_row)
{
	uint64_t acc = 0, n = 0;
	const uint64_t mask = k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
	*bad_row = -1;
	*sum = 0;
	*n_out = 0;
	for (uint64_t r = 0; r < n_rows; r++)
	{
		const char *s = flat + off[r];
		uint64_t len = off[r + 1] - off[r];
		if (orc_validate_dna(s, len) >= 0)
		{
			*bad_row = (int64_t) r;
			return ORC_INVALID_DNA;
		}
		if ((int64_t) len < (int64_t) k || k <= 0 || k > ORC_MAX_K)
		{
			*bad_row = (int64_t) r;
			return ORC_INVALID_K;
		}
		uint64_t v = 0;
		for (uint64_t p = 0; p < len; p++)
		{
			v = ((v << 2) | (uint64_t) base_code((unsigned char) s[p])) & mask;
			if (p + 1 >= (uint64_t) k)
			{
				acc += orc_mix64(v);
				n++;
			}
		}
	}
	*sum = acc;
	*n_out = n;
	return ORC_OK;
}

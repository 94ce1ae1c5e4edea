"""CPU tests: pin the oracle (C restatement + numpy restatement) against
  (1) the reference's recorded answers (tests/golden/kat.json <- kmer-tests.sql),
  (2) the reference's own code compiled unmodified (oracle/_ref) when present,
  (3) fixtures produced by (2) and committed (tests/golden/ref_vectors.json)."""
import numpy as np
import pytest

import golden_util as G
from oracle import oracle as O

KAT = G.load("kat.json")
VEC = G.load("ref_vectors.json")
SQLSTATE = {"22P02": O.SQLSTATE_22P02, "22001": O.SQLSTATE_22001, "22023": O.SQLSTATE_22023}
KAT_OP = {"eq": O.OP_EQUALS, "sw": O.OP_STARTS_WITH, "swop": O.OP_STARTS_WITH_OP, "ct": O.OP_CONTAINS}


# ----------------------------------------------------------------- (1)+(2): reference .so vs recorded psql output

def test_ref_cast_kat(ref):
    for c in KAT["cast"]:
        assert ref.cast(c["type"], c["in"]) == c["out"], c["src"]
    for c in KAT["cast_errors"]:
        with pytest.raises(O.RefSqlError) as ei:
            ref.cast(c["type"], c["in"])
        assert ei.value.sqlstate == SQLSTATE[c["sqlstate"]] and ei.value.message == c["message"] and ei.value.detail == c["detail"], c["src"]


def test_ref_generate_kat(ref):
    for c in KAT["generate"]:
        assert ref.generate_kmers(c["dna"], c["k"]) == c["out"], c["src"]
    for c in KAT["generate_errors"]:
        with pytest.raises(O.RefSqlError) as ei:
            ref.generate_kmers(c["dna"], c["k"])
        assert ei.value.sqlstate == SQLSTATE[c["sqlstate"]] and ei.value.message == c["message"], c["src"]


def test_ref_predicates_kat(ref):
    for c in KAT["predicates"]:
        assert ref.predicate(KAT_OP[c["op"]], c["a"], c["b"]) == c["out"], c
    # commutator forms: kmer <@ qkmer == qkmer @> kmer  (kmer--1.0.0.sql:156-171)
    for c in KAT["predicates"]:
        if c["op"] == "ct":
            assert ref.predicate(O.OP_CONTAINING, c["b"], c["a"]) == c["out"], c


def test_ref_group_by_kat(ref):
    for c in KAT["group_by"]:
        flat, off = O.rows_to_flat([c["dna"]])
        keys, counts, n = ref.count(flat, off, c["k"])
        assert {bytes(k).decode(): int(v) for k, v in zip(keys, counts)} == c["out"], c["src"]
    for c in KAT["count"]:
        flat, off = O.rows_to_flat([c["dna"]])
        assert ref.count(flat, off, c["k"])[2] == c["total"], c["src"]


# ----------------------------------------------------------------- oracle restatements vs recorded psql output

def _orc_pred(corc, op, a, b):
    """Evaluate a SQL-level predicate on the C oracle (a, b as in the reference's entry point)."""
    if op == O.OP_EQUALS:
        ca, la = corc.kmer_encode(a)
        return bool(corc.match_column(0, np.array([ca], np.uint64), la, b, np.array([la], np.uint8))[0])
    if op == O.OP_STARTS_WITH:      # starts_with(prefix a, kmer b)
        cb, lb = corc.kmer_encode(b)
        return bool(corc.match_column(1, np.array([cb], np.uint64), lb, a, np.array([lb], np.uint8))[0])
    if op == O.OP_STARTS_WITH_OP:   # a ^@ prefix b
        ca, la = corc.kmer_encode(a)
        return bool(corc.match_column(1, np.array([ca], np.uint64), la, b, np.array([la], np.uint8))[0])
    if op == O.OP_CONTAINS:         # contains(qkmer a, kmer b)
        cb, lb = corc.kmer_encode(b)
        return bool(corc.match_column(2, np.array([cb], np.uint64), lb, a, np.array([lb], np.uint8))[0])
    ca, la = corc.kmer_encode(a)    # containing(kmer a, qkmer b)
    return bool(corc.match_column(2, np.array([ca], np.uint64), la, b, np.array([la], np.uint8))[0])


def _np_pred(op, a, b):
    enc = lambda s: (int(O.np_encode_ascii(np.frombuffer(s.encode(), np.uint8).reshape(1, -1))[0]) if s else 0, len(s))
    if op == O.OP_EQUALS:
        c, l = enc(a); return bool(O.np_match(0, np.array([c], np.uint64), l, b)[0])
    if op == O.OP_STARTS_WITH:
        c, l = enc(b); return bool(O.np_match(1, np.array([c], np.uint64), l, a)[0])
    if op == O.OP_STARTS_WITH_OP:
        c, l = enc(a); return bool(O.np_match(1, np.array([c], np.uint64), l, b)[0])
    if op == O.OP_CONTAINS:
        c, l = enc(b); return bool(O.np_match(2, np.array([c], np.uint64), l, a)[0])
    c, l = enc(a); return bool(O.np_match(2, np.array([c], np.uint64), l, b)[0])


def test_oracle_predicates_kat(corc):
    for c in KAT["predicates"]:
        assert _orc_pred(corc, KAT_OP[c["op"]], c["a"], c["b"]) == c["out"], c
        assert _np_pred(KAT_OP[c["op"]], c["a"], c["b"]) == c["out"], c


def test_oracle_generate_and_group_kat(corc):
    for c in KAT["generate"]:
        flat, off = O.rows_to_flat([c["dna"]])
        for codes in (corc.generate(flat, off, c["k"]), O.np_generate(flat, off, c["k"])):
            assert [corc.kmer_decode(int(x), c["k"]) for x in codes] == c["out"], c["src"]
    for c in KAT["generate_errors"]:
        flat, off = O.rows_to_flat([c["dna"]])
        for f in (corc.generate, O.np_generate):
            with pytest.raises(O.OracleError) as ei:
                f(flat, off, c["k"])
            assert ei.value.code == O.ORC_INVALID_K and ei.value.row == 0
    for c in KAT["group_by"]:
        flat, off = O.rows_to_flat([c["dna"]])
        for keys, counts, n in (corc.count(flat, off, c["k"]), O.np_count(flat, off, c["k"])):
            assert {corc.kmer_decode(int(x), c["k"]): int(v) for x, v in zip(keys, counts)} == c["out"]
            assert n == sum(c["out"].values())


def test_oracle_cast_kat(corc):
    for c in KAT["cast"]:
        if c["type"] == "qkmer":
            assert corc.qkmer_parse(c["in"]) == c["out"]
        else:
            code, l = corc.kmer_encode(c["in"]) if c["type"] == "kmer" else (None, None)
            if code is not None:
                assert corc.kmer_decode(code, l) == c["out"]
    want = {("kmer", "22001"): O.ORC_KMER_TOO_LONG, ("kmer", "22P02"): O.ORC_INVALID_DNA,
            ("qkmer", "22001"): O.ORC_QKMER_TOO_LONG, ("qkmer", "22P02"): O.ORC_INVALID_QKMER}
    for c in KAT["cast_errors"]:
        if c["type"] == "dna":
            assert corc.lib.orc_validate_dna(c["in"].encode(), len(c["in"])) == c["in"].index("N")
            continue
        with pytest.raises(O.OracleError) as ei:
            (corc.kmer_encode if c["type"] == "kmer" else corc.qkmer_parse)(c["in"])
        assert ei.value.code == want[(c["type"], c["sqlstate"])]


# ----------------------------------------------------------------- (3): fixtures generated by the reference itself

@pytest.mark.parametrize("case", VEC["generate"], ids=lambda c: f"{c['input']['gen']}{c['input']['seed']}-k{c['k']}")
def test_oracle_generate_fixture(corc, case):
    flat, off = G.checked_input(case)
    a = corc.generate(flat, off, case["k"])
    b = O.np_generate(flat, off, case["k"])
    assert a.size == case["n_kmers"] and np.array_equal(a, b)
    assert G.sha(G.generate_canon_from_codes(a, case["k"])) == case["kmers_sha256"]


@pytest.mark.parametrize("case", VEC["count"], ids=lambda c: f"{c['input']['gen']}{c['input']['seed']}-k{c['k']}")
def test_oracle_count_fixture(corc, case):
    flat, off = G.checked_input(case)
    k = case["k"]
    keys, counts, n = O.np_count(flat, off, k)
    assert n == case["n_kmers"] and keys.size == case["n_distinct"] and int(counts.max()) == case["max_count"]
    assert G.sha(G.count_canon_from_codes(keys, counts, k)) == case["table_sha256"]
    if n <= 2_100_000:
        ck, cc, cn = corc.count(flat, off, k)
        assert cn == n and np.array_equal(ck, keys) and np.array_equal(cc, counts)
    if "table" in case:
        assert {corc.kmer_decode(int(x), k): int(v) for x, v in zip(keys, counts)} == case["table"]


def test_oracle_count_repetitive_fixture(corc):
    for case in VEC["count_repetitive"]:
        flat, off = O.rows_to_flat(case["rows"])
        for keys, counts, n in (corc.count(flat, off, case["k"]), O.np_count(flat, off, case["k"])):
            assert n == case["n_kmers"]
            assert {corc.kmer_decode(int(x), case["k"]): int(v) for x, v in zip(keys, counts)} == case["table"]


def test_oracle_predicates_fixture(corc):
    for c in VEC["predicates"]:
        assert _orc_pred(corc, c["op"], c["a"], c["b"]) == c["out"], c
        assert _np_pred(c["op"], c["a"], c["b"]) == c["out"], c


def test_oracle_batch_errors_fixture(corc):
    code_of = {"0x22023": O.ORC_INVALID_K, "0x22503": O.ORC_INVALID_DNA}
    for c in VEC["batch_errors"]:
        flat, off = O.rows_to_flat(c["rows"])
        for f in (corc.count, O.np_count):
            with pytest.raises(O.OracleError) as ei:
                f(flat, off, c["k"])
            assert ei.value.code == code_of[c["error"]["sqlstate"]] and ei.value.row == c["error"]["row"], c


# ----------------------------------------------------------------- live cross-check oracle vs reference .so (here only)

def test_oracle_vs_ref_live(ref, corc):
    rng = np.random.default_rng(5)
    for trial in range(6):
        rows = ["".join(rng.choice(list("ACGTacgt"), size=int(rng.integers(33, 90)))) for _ in range(150)]
        flat, off = O.rows_to_flat(rows)
        for k in (1, 4, 11, 21, 32):
            rk, rc, rn = ref.count(flat, off, k, threads=3)
            ok, oc, on = corc.count(flat, off, k)
            codes = O.np_encode_ascii(rk)
            o = np.argsort(codes)
            assert rn == on and np.array_equal(codes[o], ok) and np.array_equal(rc[o], oc)
    # column predicates
    col = rng.choice(np.frombuffer(b"ACGTacgt", np.uint8), size=(3000, 12))
    codes = O.np_encode_ascii(col)
    for op, orc_op, const in [(O.OP_EQUALS, 0, bytes(col[7]).decode()), (O.OP_STARTS_WITH, 1, bytes(col[3][:3]).decode()),
                              (O.OP_STARTS_WITH_OP, 1, ""), (O.OP_CONTAINS, 2, "NNRYNNNNSWNN"), (O.OP_CONTAINING, 2, "nnnnnnnnnnnu"),
                              (O.OP_CONTAINS, 2, "ACGT")]:
        want = ref.predicate_column(op, col.reshape(-1), 12, const)
        assert np.array_equal(corc.match_column(orc_op, codes, 12, const), want)
        assert np.array_equal(O.np_match(orc_op, codes, 12, const), want)


def test_multiset_checksum_matches_tables(corc):
    """orc_multiset_checksum (the config-scale parity check of bench.py): sum of mix64(code) over the windows of the rows
    == sum of count * mix64(code) over the GROUP BY table, for the C port, the numpy port and -- when built -- the reference."""
    from kmer_extension_b200 import datagen
    for seed, k in ((1, 5), (2, 21), (3, 31), (4, 32), (5, 1), (6, 14)):
        flat, off = datagen.synth_ragged(seed, 300, 200, min_len=33, mixed_case=True)
        s, n = corc.multiset_checksum(flat, off, k)
        keys, counts, nk = O.np_count(flat, off, k)
        assert n == nk == int(counts.sum())
        assert s == O.np_table_checksum(keys, counts)
        ck, cc, cn = corc.count(flat, off, k)
        assert s == O.np_table_checksum(ck, cc) and cn == n
    # a table with one group split in two, or one count off by one, must not pass
    flat, off = datagen.synth_reads(9, 50, 300)
    s, n = corc.multiset_checksum(flat, off, 21)
    keys, counts, _ = O.np_count(flat, off, 21)
    bad = counts.copy(); bad[0] += 1
    assert O.np_table_checksum(keys, bad) != s
    with pytest.raises(O.OracleError):
        corc.multiset_checksum(np.frombuffer(b"ACGTNACGT", np.uint8), np.array([0, 9], np.uint64), 3)

import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.load_package()
from kmer_extension_b200 import api
VEC = json.load(open("tests/golden/ref_vectors.json"))
eng = api.KmerCuda(0)
for case in VEC["count_repetitive"]:
    k = case["k"]
    for rep in range(3):
        keys, counts, n = eng.count_kmers(case["rows"], k)
        txt = [bytes(r).decode() for r in eng.decode(keys, k)]
        got = dict(zip(txt, map(int, counts)))
        want = case["table"]
        if n != case["n_kmers"] or got != want:
            print("k", k, "rep", rep, "n", n, case["n_kmers"], "distinct", len(got), len(want), "dups in out", len(txt) - len(set(txt)))
            bad = [(s, got.get(s), want.get(s)) for s in set(got) | set(want) if got.get(s) != want.get(s)]
            print("  mismatches", len(bad), bad[:6], "sum got", sum(counts.tolist()), "sum want", sum(want.values()))
        else:
            print("k", k, "rep", rep, "ok")

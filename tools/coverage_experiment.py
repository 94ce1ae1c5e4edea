"""Profiling experiment (not a bench): k=21 count over reads SAMPLED from a random genome (default 30x coverage), i.e.
repetitive input where almost every k-mer occurs ~30 times -- the opposite regime of the all-distinct bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from kmer_extension_b200 import api

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
L, K = 1000, 21
rng = np.random.default_rng(11)
G = int(n_reads * L / cov)
genome = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=G + L)]
starts = rng.integers(0, G, size=n_reads)
flat = genome[(starts[:, None] + np.arange(L)[None, :])].reshape(-1)
off = (np.arange(n_reads + 1) * L).astype(np.uint64)
d_seq = torch.from_numpy(np.concatenate([flat, np.zeros(64, np.uint8)])).cuda()
d_off = torch.from_numpy(off.astype(np.int64)).cuda()
eng = api.KmerCuda(0)
cap = eng.max_kmers(int(off[-1]), n_reads, K)
d_pairs = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
eng.set_profiling(True)
for it in range(3):
    eng.dev_count(d_seq, int(off[-1]), d_off, n_reads, K, d_pairs, algo=0)
    r = eng.dev_finish()
    print(f"coverage {cov}: n_kmers {r.n_kmers} distinct {r.n_distinct} tier2 {r.n_tier2} overflow {r.n_overflow}", eng.phases())
p = d_pairs[: r.n_distinct].cpu().numpy().view(np.uint64)
print("sum of counts", int(p[:, 1].sum()), "max count", int(p[:, 1].max()), "mean", float(p[:, 1].mean()))
assert int(p[:, 1].sum()) == r.n_kmers and np.unique(p[:, 0]).size == p.shape[0]

"""debug helper: GPU table vs oracle on a small input, lists what differs"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from kmer_extension_b200 import api, datagen
from oracle import oracle as O
eng = api.KmerCuda(0)
for k in [int(x) for x in os.environ.get("KS", "21,26,31,14").split(",")]:
    for n_rows in (300, 20000):
        flat, off = datagen.synth_reads(7 + k, n_rows, 1000)
        d_seq = torch.from_numpy(np.concatenate([flat, np.zeros(64, np.uint8)])).cuda()
        d_off = torch.from_numpy(off.astype(np.int64)).cuda()
        cap = eng.max_kmers(int(off[-1]), n_rows, k)
        d_pairs = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
        eng.dev_count(d_seq, int(off[-1]), d_off, n_rows, k, d_pairs, algo=int(os.environ.get("ALGO", "3")))
        r = eng.dev_finish()
        p = d_pairs[: r.n_distinct].cpu().numpy().view(np.uint64)
        wk, wc, wn = O.np_count(flat, off, k)
        gk, gc = p[:, 0], p[:, 1]
        o = np.argsort(gk, kind="stable"); gk, gc = gk[o], gc[o]
        dup = int((gk[1:] == gk[:-1]).sum())
        missing = np.setdiff1d(wk, gk).size
        extra = np.setdiff1d(gk, wk).size
        print(f"k={k} rows={n_rows}: n_kmers {r.n_kmers}/{wn} groups {gk.size}/{wk.size} dup_keys {dup} missing {missing} extra {extra} "
              f"sum_counts {int(gc.sum())} tier2 {r.n_tier2} overflow {r.n_overflow}", flush=True)

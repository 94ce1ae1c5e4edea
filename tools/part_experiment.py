"""Profiling experiment (not a bench): time the partition kernel with record stores and/or slot atomics removed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from kmer_extension_b200 import api, datagen

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
flat, off = datagen.synth_reads(2, n_rows, 1000)
d_seq = torch.from_numpy(np.concatenate([flat, np.zeros(64, np.uint8)])).cuda()
d_off = torch.from_numpy(off.astype(np.int64)).cuda()
eng = api.KmerCuda(0, os.environ.get('KMER_CUDA_LIB', api.LIB_PATH))
KK = int(os.environ.get("KMER_K", "21"))
cap = eng.max_kmers(int(off[-1]), n_rows, KK)
d_pairs = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
eng.set_profiling(True)
for it in range(3):
    try:
        eng.dev_count(d_seq, int(off[-1]), d_off, n_rows, KK, d_pairs, algo=int(os.environ.get('KMER_ALGO','3')))
        r = eng.dev_finish()
        extra = f"tier2_kmers={r.n_tier2} recounted={r.n_overflow} distinct={r.n_distinct}"
    except api.KmerSqlError as e:
        extra = f"error: {e}"
    print(f"k={KK}", [(n, round(ms, 3)) for n, ms in eng.phases()], extra)

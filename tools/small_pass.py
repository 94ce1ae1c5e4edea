"""Small pass over every kernel of the counting / matching path, checked against the oracle (compute-sanitizer is closed on this GPU pool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import __graft_entry__ as g
g.load_package()
from kmer_extension_b200 import api, datagen
from oracle import oracle as O

eng = api.KmerCuda(0)
flat, off = datagen.synth_reads(5, 400, 1000)
rep = O.rows_to_flat(["ACGTTGCA" * 600, "A" * 9000, "t" * 3000])
flat2 = np.concatenate([flat, np.frombuffer(rep[0], np.uint8)])
off2 = np.concatenate([off, off[-1] + rep[1][1:]])
for k in (5, 17, 21, 27, 32):
    ok, oc, on = O.np_count(flat2, off2, k)
    keys, counts, n = eng.count_kmers(flat2, k, off2)
    o = np.argsort(keys)
    assert n == on and np.array_equal(keys[o], ok) and np.array_equal(counts[o], oc), k
    u, keys, counts, n = eng.count_kmers_split(flat2, k, off2)
    allk = np.concatenate([u, keys]); allc = np.concatenate([np.ones(u.size, np.uint64), counts]); o = np.argsort(allk)
    assert np.array_equal(allk[o], ok) and np.array_equal(allc[o], oc), k
    codes = eng.generate_kmers(flat2, k, off2)
    assert np.array_equal(codes, O.np_generate(flat2, off2, k))
# sharded, two simulated ranks
k, G = 21, 2
cut = len(off2) // 2
shards = [(flat2[: int(off2[cut])], off2[: cut + 1]), (flat2[int(off2[cut]):], off2[cut:] - off2[cut])]
ok, oc, on = O.np_count(flat2, off2, k)
plan = eng.shard_plan(on, k, G)
sr, sf = [], []
for f, o_ in shards:
    d_seq = torch.from_numpy(np.concatenate([f, np.zeros(64, np.uint8)])).cuda()
    d_off = torch.from_numpy(o_.astype(np.int64)).cuda()
    recs = torch.empty(plan.recs_bytes_per_peer * G, dtype=torch.uint8, device="cuda")
    fill = torch.empty(plan.buckets_per_rank * G, dtype=torch.int64, device="cuda")
    eng.dev_shard_partition(d_seq, int(o_[-1]), d_off, len(o_) - 1, plan, recs, fill)
    eng.dev_finish()
    sr.append(recs); sf.append(fill)
ka, ca = [], []
rb, fb = plan.recs_bytes_per_peer, plan.buckets_per_rank
for owner in range(G):
    rr = torch.cat([sr[r][owner * rb:(owner + 1) * rb] for r in range(G)])
    rf = torch.cat([sf[r][owner * fb:(owner + 1) * fb] for r in range(G)])
    pairs = torch.empty((on + 16, 2), dtype=torch.int64, device="cuda")
    eng.dev_shard_count(plan, rr, rf, pairs)
    r = eng.dev_finish()
    pp = pairs[: r.n_distinct].cpu().numpy().view(np.uint64)
    ka.append(pp[:, 0].copy()); ca.append(pp[:, 1].copy())
keys = np.concatenate(ka); counts = np.concatenate(ca); o = np.argsort(keys)
assert np.array_equal(keys[o], ok) and np.array_equal(counts[o], oc)
# matching: few constants (pairwise kernel) and many (table kernel)
col = datagen.synth_kmer_codes(7, 3000, 12)
pats = datagen.synth_qkmers(8, 130, 12, with_n=True)
for ps in (pats[:5], pats):
    bits, hits = eng.match(api.OP_CONTAINS, col, 12, ps)
    for i, p in enumerate(ps):
        want = O.np_match(2, col, 12, p).astype(bool)
        assert np.array_equal(bits[i], want) and hits[i] == want.sum()
print("sanitize workload ok")

"""Worker of tests/test_multigpu_nccl.py: launched by torchrun, one rank per GPU, real NCCL.
Every rank counts ITS rows with ShardedCounter; the union of the per-rank tables must be exactly the oracle's
GROUP BY table of ALL rows (no group lost, none split over two owners)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import conftest  # noqa: F401  (imports the package)
from kmer_extension_b200 import api, datagen, sharded
from oracle import oracle as O


def gather_tables(keys, counts, world, rank):
    """every rank's (keys, counts) on rank 0, concatenated"""
    n = torch.tensor([keys.numel()], dtype=torch.int64, device="cuda")
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    m = int(max(int(x.item()) for x in ns))
    pad_k = torch.zeros(m, dtype=torch.int64, device="cuda"); pad_k[:keys.numel()] = keys
    pad_c = torch.zeros(m, dtype=torch.int64, device="cuda"); pad_c[:counts.numel()] = counts
    ks = [torch.empty_like(pad_k) for _ in range(world)]
    cs = [torch.empty_like(pad_c) for _ in range(world)]
    dist.all_gather(ks, pad_k)
    dist.all_gather(cs, pad_c)
    K = np.concatenate([ks[r][:int(ns[r].item())].cpu().numpy().view(np.uint64) for r in range(world)])
    Cn = np.concatenate([cs[r][:int(ns[r].item())].cpu().numpy().view(np.uint64) for r in range(world)])
    return K, Cn


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = api.KmerCuda(local)
    sh = sharded.ShardedCounter(eng)
    cases = []
    for k in (21, 31, 14, 5):
        flat, off = datagen.synth_reads(100 + rank, 3000 + 500 * rank, 700)     # ragged over the ranks
        cases.append((f"random k={k}", k, flat, off))
    # skewed input: poly-A / poly-T runs and one read repeated many times on every rank
    rng = np.random.default_rng(7 + rank)
    rows = []
    rep = "".join(rng.choice(list("ACGT"), 400))
    for i in range(2500):
        r = i % 5
        if r == 0: rows.append("A" * int(rng.integers(40, 300)))
        elif r == 1: rows.append(rep)
        elif r == 2: rows.append("T" * 64 + "".join(rng.choice(list("ACGT"), 200)))
        else: rows.append("".join(rng.choice(list("ACGT"), int(rng.integers(40, 500)))))
    fs, os_ = O.rows_to_flat(rows)
    for k in (21, 32):
        cases.append((f"skewed k={k}", k, np.frombuffer(fs, dtype=np.uint8).copy(), os_))
    ok = True
    for name, k, flat, off in cases:
        n_rows, n_bases = len(off) - 1, int(off[-1])
        d_seq = torch.from_numpy(np.concatenate([np.asarray(flat, dtype=np.uint8), np.zeros(64, np.uint8)])).cuda()
        d_off = torch.from_numpy(np.asarray(off).astype(np.int64)).cuda()
        cap = eng.max_kmers(n_bases, n_rows, k) * world + 4096
        d_pairs = torch.empty((cap, 2), dtype=torch.int64, device="cuda")
        nd, nk, info = sh.count(d_seq, n_bases, d_off, n_rows, k, d_pairs)
        K, Cn = gather_tables(d_pairs[:nd, 0].contiguous(), d_pairs[:nd, 1].contiguous(), world, rank)
        # all rows of all ranks, gathered on every rank for the oracle
        blobs = [None] * world
        dist.all_gather_object(blobs, (np.asarray(flat, dtype=np.uint8).tobytes(), np.asarray(off, dtype=np.uint64).tolist()))
        if rank == 0:
            allflat = np.concatenate([np.frombuffer(b, dtype=np.uint8) for b, _ in blobs])
            alloff, base = [0], 0
            for b, o in blobs:
                alloff += [base + int(x) for x in o[1:]]
                base += o[-1]
            wk, wc, wn = O.np_count(allflat, np.array(alloff, dtype=np.uint64), k)
            o = np.argsort(K, kind="stable")
            good = K.size == wk.size and np.array_equal(K[o], wk) and np.array_equal(Cn[o], wc)
            print(f"[nccl x{world}] {name}: {'ok' if good else 'MISMATCH'} groups={K.size}/{wk.size} tier2={info.get('tier2_kmers')} fallback={info.get('fallback')}", flush=True)
            ok = ok and good
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
    dist.broadcast(flag, 0)
    print(f"[nccl x{world}] rank {rank} exchange={'pull' if sh._pull is not None else 'nccl'}", flush=True)
    sh.close()
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

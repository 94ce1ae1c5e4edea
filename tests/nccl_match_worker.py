"""Worker of tests/test_sharded_match_gpu.py::test_sharded_match_over_nccl_2gpu: launched by torchrun, one rank per GPU, real
NCCL.  Every rank matches ITS slice of one k-mer column against the replicated constants with ShardedMatcher; the ranks' words
concatenate to the oracle's bit matrix of the whole column and the all-reduced hit counts are the oracle's row sums."""
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


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = api.KmerCuda(local)
    sm = sharded.ShardedMatcher(eng)
    ok = True
    cases = [("contains k=12, 40 patterns", api.OP_CONTAINS, 12, 300_007, datagen.synth_qkmers(71, 39, 12, with_n=True) + ["n" * 12], None),
             ("contains k=12, 200 patterns (table kernel)", api.OP_CONTAINS, 12, 100_003, datagen.synth_qkmers(72, 200, 12, with_n=True), None),
             ("equals + starts_with k=32", api.OP_EQUALS, 32, 200_001, None, [api.OP_EQUALS, api.OP_STARTS_WITH])]
    for name, op, k, m, consts, ops in cases:
        col = datagen.synth_kmer_codes(80 + k, m, k)                 # the same column on every rank; a rank uploads its slice only
        if consts is None:
            txt = bytes(O.np_decode(col[:1], k)[0]).decode()
            consts = [txt, txt[:6]]
        lo, hi = sm.slice_of(m, rank, world)
        d_codes = torch.from_numpy(col[lo:hi].view(np.int64).copy()).cuda()
        wl = sm.words_per_row(hi - lo)
        d_bits = torch.zeros(max(len(consts) * wl, 1), dtype=torch.int32, device="cuda")
        d_hits = torch.zeros(len(consts), dtype=torch.int64, device="cuda")
        sm.match(op, d_codes, hi - lo, k, consts, d_bits, d_hits, ops=ops)
        full = sm.gather_bits(d_bits, hi - lo, m, len(consts)).cpu().numpy()
        got = np.unpackbits(full.view(np.uint32).view(np.uint8), axis=1, bitorder="little")[:, :m].astype(bool)
        hits = d_hits.cpu().numpy()
        good = True
        for i, c in enumerate(consts):
            want = O.np_match(op if ops is None else ops[i], col, k, c).astype(bool)
            good = good and np.array_equal(got[i], want) and int(hits[i]) == int(want.sum())
        print(f"[nccl match x{world}] rank {rank} {name}: {'ok' if good else 'MISMATCH'} slice=[{lo},{hi})", flush=True)
        ok = ok and good
    flag = torch.tensor([1 if ok else 0], dtype=torch.int64, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    eng.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()

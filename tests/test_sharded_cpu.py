"""CPU (gloo, world_size 2) test of the sharded-count host logic: bucket ownership, equal-split all-to-all
layout, dense all-reduce path and error agreement.  The CUDA engine cannot run here, so a numpy stand-in
engine with the same method surface produces/consumes the opaque record buffers; the oracle is the checker."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


class _Plan:
    pass


class _Res:
    def __init__(self, n_kmers=0, n_distinct=0):
        self.n_kmers, self.n_distinct, self.n_tier2, self.n_overflow = n_kmers, n_distinct, 0, 0


class StandInEngine:
    """Same surface as KmerCuda's sharding calls; a 'record' is one raw k-mer code (8 bytes, L = 1)."""

    def __init__(self, O):
        self.O = O
        self._res = _Res()

    def max_kmers(self, n_bases, n_rows, k):
        return max(n_bases - n_rows * (k - 1), 0)

    def shard_plan(self, total, k, n_ranks):
        p = _Plan()
        p.n_ranks, p.k = n_ranks, k
        p.buckets_per_rank = max(1, (total // 50 + n_ranks - 1) // n_ranks)
        p.n_buckets = p.buckets_per_rank * n_ranks
        p.cap = 400
        p.rec_bytes = 8
        p.recs_bytes_per_peer = p.buckets_per_rank * p.cap * 8
        p.fill_bytes_per_peer = p.buckets_per_rank * 8
        return p

    @staticmethod
    def _bucket(codes, nb):
        h = (codes * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)
        return (h % np.uint64(nb)).astype(np.int64)

    def dev_shard_partition(self, d_seq, n_bases, d_off, n_rows, plan, send_recs, send_fill, stream=None):
        flat = d_seq.numpy()[:n_bases]
        off = d_off.numpy().astype(np.uint64)
        codes = self.O.np_generate(flat, off, plan.k)   # raises OracleError on bad input
        b = self._bucket(codes, plan.n_buckets)
        recs = send_recs.numpy().view(np.uint64).reshape(plan.n_buckets, plan.cap)
        fill = send_fill.numpy()
        fill[:] = 0
        for code, bb in zip(codes, b):
            s = fill[bb] & 0xFFFFFFFF
            assert s < plan.cap
            recs[bb, s] = code
            fill[bb] += (1 << 32) | 1
        self._res = _Res(n_kmers=codes.size)

    def dev_shard_count(self, plan, recv_recs, recv_fill, d_pairs, stream=None):
        recs = recv_recs.numpy().view(np.uint64).reshape(plan.n_ranks, plan.buckets_per_rank, plan.cap)
        fill = recv_fill.numpy().reshape(plan.n_ranks, plan.buckets_per_rank)
        got = [recs[s, b, : int(fill[s, b] & 0xFFFFFFFF)] for s in range(plan.n_ranks) for b in range(plan.buckets_per_rank)]
        allc = np.concatenate(got) if got else np.zeros(0, np.uint64)
        keys, counts = np.unique(allc, return_counts=True)
        out = d_pairs.numpy().view(np.uint64)
        out[: keys.size, 0] = keys
        out[: keys.size, 1] = counts.astype(np.uint64)
        self._res = _Res(n_kmers=int(allc.size), n_distinct=int(keys.size))

    def dev_dense_table(self, d_seq, n_bases, d_off, n_rows, k, table, stream=None):
        codes = self.O.np_generate(d_seq.numpy()[:n_bases], d_off.numpy().astype(np.uint64), k)
        t = table.numpy()
        t[:] = 0
        np.add.at(t, codes.astype(np.int64), 1)
        self._res = _Res(n_kmers=codes.size)

    def dev_dense_emit(self, table, k, rank, n_ranks, d_pairs, stream=None):
        t = table.numpy()
        bins = np.nonzero(t)[0]
        bins = bins[bins % n_ranks == rank]
        out = d_pairs.numpy().view(np.uint64)
        out[: bins.size, 0] = bins.astype(np.uint64)
        out[: bins.size, 1] = t[bins].astype(np.uint64)
        self._res = _Res(n_kmers=int(t[bins].sum()), n_distinct=int(bins.size))

    def dev_match(self, op, d_codes, m, k, consts, d_bits, d_hits, d_lens=None, ops=None, stream=None):
        """same bit layout as kmer_cuda_dev_match: row c = words_per_row int32 words, bit i%32 of word i/32 = k-mer i"""
        co = self.O.COracle()
        codes = d_codes.numpy().view(np.uint64)[:m]
        wpr = (m + 31) // 32
        bits = d_bits.numpy().view(np.uint32)[: len(consts) * wpr].reshape(len(consts), wpr)
        for c, text in enumerate(consts):
            if text == "boom":
                raise ValueError("bad constant")
            o = op if ops is None else int(ops[c])
            b = co.match_column(o, codes, k, text.lower()).astype(np.uint8)
            row = np.packbits(np.concatenate([b, np.zeros(wpr * 32 - m, np.uint8)]), bitorder="little").view(np.uint32)
            bits[c] = row
            d_hits[c] = int(b.sum())
        self._res = _Res()

    def dev_finish(self, stream=None):
        return self._res


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest  # noqa: F401
    from kmer_extension_b200 import datagen, sharded
    from oracle import oracle as O
    try:
        eng = StandInEngine(O)
        sc = sharded.ShardedCounter(eng, device=torch.device("cpu"))
        results = {}
        for k in (5, 15, 21):
            flat, off = datagen.synth_ragged(100 + rank, 60, 90, min_len=32, mixed_case=True)   # this rank's rows
            d_seq = torch.from_numpy(flat.copy())
            d_off = torch.from_numpy(off.astype(np.int64))
            pairs = torch.zeros((int(off[-1]) * world + 1024, 2), dtype=torch.int64)
            nd, nk, info = sc.count(d_seq, int(off[-1]), d_off, len(off) - 1, k, pairs)
            p = pairs[:nd].numpy().view(np.uint64)
            results[k] = (p[:, 0].copy(), p[:, 1].copy(), nk)
        # error agreement: rank 1 has a bad row, both ranks must raise
        rows = ["ACGTACGTACGTACGTACGTAC", "ACGTNACGTACGTACGTACGTA" if rank == 1 else "ACGTTACGTACGTACGTACGTA"]
        fb, off = O.rows_to_flat(rows)
        raised = False
        try:
            sc.count(torch.from_numpy(np.frombuffer(fb, np.uint8).copy()), int(off[-1]), torch.from_numpy(off.astype(np.int64)),
                     2, 15, torch.zeros((100, 2), dtype=torch.int64))
        except Exception:
            raised = True
        np.savez(Path(out_dir) / f"rank{rank}.npz", raised=raised,
                 **{f"k{k}_{n}": v for k, (a, b, c) in results.items() for n, v in (("keys", a), ("counts", b), ("nk", np.array(c)))})
    finally:
        dist.destroy_process_group()


def test_sharded_count_two_ranks_gloo(tmp_path):
    import conftest  # noqa: F401
    from kmer_extension_b200 import datagen
    from oracle import oracle as O
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert all(bool(o["raised"]) for o in outs), "an input error on one rank must abort every rank"
    # the union of the per-rank tables must equal the oracle's count over ALL rows, and be disjoint
    for k in (5, 15, 21):
        all_flat, all_off = [], [np.zeros(1, np.uint64)]
        for r in range(world):
            flat, off = datagen.synth_ragged(100 + r, 60, 90, min_len=32, mixed_case=True)
            all_flat.append(flat)
            all_off.append(off[1:] + all_off[-1][-1])
        flat = np.concatenate(all_flat)
        off = np.concatenate(all_off)
        ok, oc, on = O.np_count(flat, off, k)
        keys = np.concatenate([o[f"k{k}_keys"] for o in outs])
        counts = np.concatenate([o[f"k{k}_counts"] for o in outs])
        assert np.unique(keys).size == keys.size, "ranks reported the same k-mer twice"
        order = np.argsort(keys)
        assert np.array_equal(keys[order], ok) and np.array_equal(counts[order], oc)
        assert sum(int(o[f"k{k}_nk"]) for o in outs) == on


def _match_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import conftest  # noqa: F401
    from kmer_extension_b200 import datagen, sharded
    from oracle import oracle as O
    try:
        sm = sharded.ShardedMatcher(StandInEngine(O), device=torch.device("cpu"))
        m_total, k = 1000 + 37, 12                                   # not a multiple of 32: the last rank's row ends mid-word
        col = datagen.synth_kmer_codes(44, m_total, k)
        consts = datagen.synth_qkmers(45, 7, k, with_n=True) + ["n" * k]
        lo, hi = sm.slice_of(m_total, rank, world)
        assert lo % 32 == 0 and (hi % 32 == 0 or hi == m_total)
        d_codes = torch.from_numpy(col[lo:hi].view(np.int64).copy())
        wl = sm.words_per_row(hi - lo)
        d_bits = torch.zeros(len(consts) * wl, dtype=torch.int32)
        d_hits = torch.zeros(len(consts), dtype=torch.int64)
        sm.match(2, d_codes, hi - lo, k, consts, d_bits, d_hits)
        full = sm.gather_bits(d_bits, hi - lo, m_total, len(consts))
        hits = d_hits.numpy().copy()
        # one rank fails alone (a device error would look like this): every rank must raise, nobody may hang
        raised = False
        try:
            sm.match(2, d_codes, hi - lo, k, ["boom"] if rank == 1 else ["acgtacgtacgt"], d_bits, d_hits)
        except Exception:
            raised = True
        np.savez(Path(out_dir) / f"m{rank}.npz", bits=full.numpy(), hits=hits, lo=lo, hi=hi, raised=raised)
    finally:
        dist.destroy_process_group()


def test_sharded_match_two_ranks_gloo(tmp_path):
    """ShardedMatcher: the column split on 32-k-mer boundaries, constants replicated, hit counts all-reduced; the ranks' words
    concatenate to the oracle's bit matrix of the whole column."""
    import conftest  # noqa: F401
    from kmer_extension_b200 import datagen
    from oracle import oracle as O
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_match_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"m{r}.npz") for r in range(world)]
    m_total, k = 1037, 12
    col = datagen.synth_kmer_codes(44, m_total, k)
    consts = datagen.synth_qkmers(45, 7, k, with_n=True) + ["n" * k]
    co = O.COracle()
    want = np.stack([co.match_column(2, col, k, c.lower()).astype(bool) for c in consts])
    assert int(outs[0]["hi"]) == int(outs[1]["lo"]) and int(outs[1]["hi"]) == m_total and int(outs[0]["lo"]) == 0
    for o in outs:
        assert bool(o["raised"]), "a failure on one rank must raise on every rank"
        got = np.unpackbits(o["bits"].view(np.uint32).view(np.uint8), axis=1, bitorder="little")[:, :m_total].astype(bool)
        assert np.array_equal(got, want)
        assert np.array_equal(o["hits"].astype(np.int64), want.sum(axis=1).astype(np.int64))
    assert want[-1].all() and want[:-1].sum() > 0

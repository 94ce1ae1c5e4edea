"""The device-side seeded generator (csrc/synth.cu, kmer_cuda_dev_synth_reads; SURVEY 8d2: data_generator.py:4-11 restated,
seeded, shape-parameterised) against its numpy restatement (datagen.synth_reads_counter) and, counted in place, against the
oracle.  CPU part: the restatement against a plain-Python statement of the same formula, shard independence, uniformity."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "kmer-extension_b200"
M64 = (1 << 64) - 1


def _py_table(seed, n, g0=0):
    def sm(z):
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)
    return np.frombuffer(bytes(b"ACGT"[(sm((seed + (((g0 + i) >> 5) + 1) * 0x9E3779B97F4A7C15) & M64) >> (2 * ((g0 + i) & 31))) & 3]
                               for i in range(n)), dtype=np.uint8)


def build_c_test() -> Path:
    from oracle import oracle as O
    O.build(ref=False)
    exe = ROOT / "tests" / "c" / "test_synth"
    cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-I", str(ROOT / "include"), "-I", "/usr/local/cuda/include",
           str(ROOT / "tests" / "c" / "test_synth.c"), "-o", str(exe), f"-L{PKG}", "-lkmer_cuda", f"-L{ROOT / 'oracle'}", "-lkmer_oracle",
           "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{ROOT / 'oracle'}"]
    subprocess.run(cmd, check=True)
    return exe


def test_counter_table_restatement():
    import conftest  # noqa: F401
    from kmer_extension_b200 import datagen
    for seed, rows, L, first in ((3, 7, 101, 5), (M64, 3, 64, 0), (0, 1, 1, 10**9), (2, 0, 50, 3)):
        flat, off = datagen.synth_reads_counter(seed, rows, L, first_row=first)
        assert np.array_equal(flat, _py_table(seed, rows * L, first * L))
        assert np.array_equal(off, np.arange(rows + 1, dtype=np.uint64) * np.uint64(L))
    # a rank's row range of the table == the same rows of the whole table, wherever the range starts
    whole, _ = datagen.synth_reads_counter(9, 40, 77)
    for first, n in ((0, 40), (13, 20), (39, 1)):
        part, _ = datagen.synth_reads_counter(9, n, 77, first_row=first)
        assert np.array_equal(part, whole[first * 77:(first + n) * 77])
    # i.i.d. uniform over ACGT (data_generator.py:4): single bases and adjacent pairs, chi-square well inside 5 sigma
    big, _ = datagen.synth_reads_counter(11, 2000, 1000)
    assert set(np.unique(big)) == set(b"ACGT")
    lut = np.zeros(256, np.int64); lut[list(b"ACGT")] = [0, 1, 2, 3]
    b = lut[big]
    c1 = np.bincount(b, minlength=4)
    assert (((c1 - b.size / 4) ** 2) / (b.size / 4)).sum() < 30
    c2 = np.bincount(b[:-1] * 4 + b[1:], minlength=16)
    assert (((c2 - (b.size - 1) / 16) ** 2) / ((b.size - 1) / 16)).sum() < 60
    # different seeds give different tables
    other, _ = datagen.synth_reads_counter(12, 2000, 1000)
    assert 0.70 < (other != big).mean() < 0.80


def test_c_synth_host_program_builds():
    assert build_c_test().exists()


@pytest.mark.gpu
def test_device_generator_c_host_vs_restatement_and_oracle():
    exe = build_c_test()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_device_generator_vs_numpy():
    import torch
    import conftest  # noqa: F401
    from kmer_extension_b200 import api, datagen
    from oracle import oracle as O
    eng = api.KmerCuda(0)
    try:
        for seed, first, rows, L in ((2, 0, 5000, 1000), (3, 777, 1234, 999), (5, 10**7, 17, 33), (6, 0, 1, 1)):
            n = rows * L
            d_seq = torch.zeros(((n + 15) & ~15) + 64, dtype=torch.uint8, device="cuda")
            d_off = torch.zeros(rows + 1, dtype=torch.int64, device="cuda")
            eng.dev_synth_reads(seed, first, rows, L, d_seq, d_off)
            eng.dev_finish()
            flat, off = datagen.synth_reads_counter(seed, rows, L, first_row=first)
            assert np.array_equal(d_seq[:n].cpu().numpy(), flat), (seed, first, rows, L)
            assert not d_seq[n:].any().item(), "the padding behind the column must stay untouched"
            assert np.array_equal(d_off.cpu().numpy().astype(np.uint64), off)
        # the generated column counted in place == the oracle over the restated rows
        rows, L, k = 4000, 500, 21
        n = rows * L
        d_seq = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
        d_off = torch.zeros(rows + 1, dtype=torch.int64, device="cuda")
        eng.dev_synth_reads(42, 0, rows, L, d_seq, d_off)
        d_pairs = torch.empty((eng.max_kmers(n, rows, k), 2), dtype=torch.int64, device="cuda")
        eng.dev_count(d_seq, n, d_off, rows, k, d_pairs)
        res = eng.dev_finish()
        flat, off = datagen.synth_reads_counter(42, rows, L)
        wk, wc, wn = O.np_count(flat, off, k)
        got = d_pairs[:res.n_distinct].cpu().numpy().view(np.uint64)
        o = np.argsort(got[:, 0])
        assert res.n_kmers == wn and np.array_equal(got[o, 0], wk) and np.array_equal(got[o, 1], wc)
    finally:
        eng.close()

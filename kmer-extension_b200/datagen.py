"""Seeded synthetic inputs for the k-mer hot path.

Restates the distribution of the reference's data_generator.py (data_generator.py:4-11: characters
drawn i.i.d. uniform with ``random.choices`` over ACGT for dna/kmer and over the 14-letter set
ACGTRYKMSWBDHV for qkmer; lengths uniform on [1, max]) with two changes the benchmark configs need:
it is seeded, and the shape (row count, read length) is a parameter -- the shipped script is
unseeded and fixed at 1 000 rows of at most 50 bases (data_generator.py:15-20).
Text is upper-case like the script's; the reference folds case on input (kmer.c:28-29).
"""
from __future__ import annotations

import numpy as np

DNA_CHARS = np.frombuffer(b"ACGT", dtype=np.uint8)            # data_generator.py:4
QKMER_CHARS = np.frombuffer(b"ACGTRYKMSWBDHV", dtype=np.uint8)  # data_generator.py:6 (no N, no U)


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def synth_reads(seed: int, n_rows: int, read_len: int, chars: np.ndarray = DNA_CHARS):
    """n_rows fixed-length reads -> (flat uint8[n_rows*read_len], offsets uint64[n_rows+1])."""
    rng = _rng(seed)
    n = n_rows * read_len
    flat = np.empty(n, dtype=np.uint8)
    step = 1 << 26
    for lo in range(0, n, step):  # chunked so that 1e9 bases do not need an 8 GB index temporary
        hi = min(n, lo + step)
        flat[lo:hi] = chars[rng.integers(0, len(chars), size=hi - lo, dtype=np.uint8)]
    off = np.arange(n_rows + 1, dtype=np.uint64) * np.uint64(read_len)
    return flat, off


def synth_ragged(seed: int, n_rows: int, max_len: int, min_len: int = 1, chars: np.ndarray = DNA_CHARS,
                 mixed_case: bool = False):
    """Rows with length uniform on [min_len, max_len] (data_generator.py:9-10)."""
    rng = _rng(seed)
    lens = rng.integers(min_len, max_len + 1, size=n_rows, dtype=np.int64)
    off = np.zeros(n_rows + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens).astype(np.uint64)
    flat = chars[rng.integers(0, len(chars), size=int(off[-1]), dtype=np.uint8)]
    if mixed_case:
        flat = flat | (rng.integers(0, 2, size=flat.size, dtype=np.uint8) << 5).astype(np.uint8)
    return flat, off


def synth_kmer_codes(seed: int, m: int, k: int) -> np.ndarray:
    """m uniform k-mers as packed codes (a=0 c=1 g=2 t=3, first base most significant)."""
    rng = _rng(seed)
    if k == 0:
        return np.zeros(m, dtype=np.uint64)
    if k == 32:
        return rng.integers(0, 1 << 64, size=m, dtype=np.uint64)
    return rng.integers(0, 1 << (2 * k), size=m, dtype=np.uint64)


def synth_qkmers(seed: int, p: int, k: int, with_n: bool = False) -> list[str]:
    """p IUPAC patterns of length k over the script's 14 letters (optionally + N)."""
    rng = _rng(seed)
    chars = QKMER_CHARS if not with_n else np.concatenate([QKMER_CHARS, np.frombuffer(b"N", np.uint8)])
    a = chars[rng.integers(0, len(chars), size=(p, k))]
    return [bytes(r).decode() for r in a]


# --------------------------------------------------------------------------- counter-based table (the device-side generator)
_GOLD = np.uint64(0x9E3779B97F4A7C15)
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _splitmix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_reads_counter(seed: int, n_rows: int, read_len: int, first_row: int = 0):
    """Rows [first_row, first_row + n_rows) of the table kmer_cuda_dev_synth_reads writes (csrc/synth.cu), bit for bit:
    base g of the whole table = "ACGT"[(splitmix64(seed + (g//32 + 1) * 0x9E3779B97F4A7C15) >> 2*(g%32)) & 3].
    Same distribution as synth_reads (data_generator.py:4-11: i.i.d. uniform ACGT, upper case), but every base depends on
    (seed, position) only, so a rank of a sharded run generates its own row range of ONE table."""
    n = n_rows * read_len
    g0 = first_row * read_len
    flat = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        g = np.arange(g0 + lo, g0 + hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            w = _splitmix64(np.uint64(seed & ((1 << 64) - 1)) + ((g >> np.uint64(5)) + np.uint64(1)) * _GOLD)
        flat[lo:hi] = _ACGT[((w >> (np.uint64(2) * (g & np.uint64(31)))) & np.uint64(3)).astype(np.int64)]
    off = np.arange(n_rows + 1, dtype=np.uint64) * np.uint64(read_len)
    return flat, off

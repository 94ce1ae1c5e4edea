#!/usr/bin/env python
"""Writes the bench workload (kmer-extension_b200/datagen.py: i.i.d. uniform ACGT reads, the distribution of the reference's
data_generator.py:4-11, seeded) as a one-column CSV for `\\copy reads FROM ...` -- the input of baseline.sql.

    python tools/pg_baseline/make_reads_csv.py 1000000 reads_1g.csv     # configs[1]: 10^6 reads x 1000 bases, seed 2
"""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
spec = importlib.util.spec_from_file_location("datagen", ROOT / "kmer-extension_b200" / "datagen.py")
datagen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(datagen)

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
out = sys.argv[2] if len(sys.argv) > 2 else "reads.csv"
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 2
flat, off = datagen.synth_reads(seed, n_reads, 1000)
with open(out, "wb") as f:
    for i in range(n_reads):
        f.write(bytes(flat[int(off[i]):int(off[i + 1])]))
        f.write(b"\n")
print(f"{n_reads} reads of 1000 bases (seed {seed}) -> {out}")

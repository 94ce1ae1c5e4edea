"""kmer-extension_b200: B200-native (sm_100a) implementation of the kmer extension's data-parallel
hot path -- generate_kmers, GROUP BY k-mer counting, batched equals/starts_with/contains -- behind
a C-ABI library (libkmer_cuda.so, include/kmer_cuda.h).

The directory name carries a hyphen (it mirrors the reference repository's name), so it is loaded
through ``importlib`` under the module name ``kmer_extension_b200``; see ``tests/conftest.py`` and
``__graft_entry__.load_package()``.
"""

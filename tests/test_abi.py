"""CPU tests of the boundary: libkmer_cuda.so loads, exports every symbol include/kmer_cuda.h
declares, and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

import conftest  # noqa: F401
from kmer_extension_b200 import api

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    if not api.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return api.load_library()


def test_header_symbols_exported(lib):
    header = (ROOT / "include" / "kmer_cuda.h").read_text()
    declared = sorted(set(re.findall(r"\b(kmer_cuda_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/kmer_cuda.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == declared


def test_abi_version_and_helpers(lib):
    assert lib.kmer_cuda_abi_version() == 2
    assert lib.kmer_cuda_max_kmers(1000, 1, 21) == 980
    assert lib.kmer_cuda_max_kmers(10, 5, 21) == 0
    assert lib.kmer_cuda_max_kmers(10, 1, 0) == 0
    assert lib.kmer_cuda_max_kmers(10, 1, 33) == 0


def test_struct_layouts_match_header():
    assert C.sizeof(api.KmerCudaError) == 4 + 6 + 160 + 160 + 6 + 8  # int, char[6], 2*char[160], pad to 8, int64
    assert C.sizeof(api.KmerDevResult) == 40
    assert C.sizeof(api.KmerShardPlan) == 72


def test_no_cpu_fallback_without_gpu(lib):
    """Without a device the product path must fail loudly, never compute on the CPU."""
    if lib.kmer_cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.KmerSqlError) as ei:
        api.KmerCuda(0)
    assert ei.value.status == api.KMER_ERR_NO_DEVICE
    assert "no CPU path" in ei.value.message


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    for p in (ROOT / "kmer-extension_b200").rglob("*"):
        if p.suffix in {".py", ".cu", ".cuh", ".c", ".h"}:
            txt = p.read_text()
            assert "oracle" not in txt.lower().replace("oracle/pgshim", ""), f"{p} mentions the oracle"

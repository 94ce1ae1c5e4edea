"""Sharded matching (shard the k-mer column, replicate the constants, no collective) against the oracle:
  * kmer_cuda_multi_submit_match from a plain C host (tests/c/test_multi_match.c): two contexts on GPU 0 always, two real GPUs
    when the box has them, one device as the degenerate case;
  * ShardedMatcher (kmer-extension_b200/sharded.py) on one GPU without a process group, and over real NCCL on 2 GPUs.
The CPU plumbing test (gloo, world_size 2) is tests/test_sharded_cpu.py::test_sharded_match_two_ranks_gloo."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "kmer-extension_b200"


def build_c_test() -> Path:
    from oracle import oracle as O
    O.build(ref=False)
    exe = ROOT / "tests" / "c" / "test_multi_match"
    cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "test_multi_match.c"), "-o", str(exe),
           f"-L{PKG}", "-lkmer_cuda", f"-L{ROOT / 'oracle'}", "-lkmer_oracle", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{ROOT / 'oracle'}"]
    subprocess.run(cmd, check=True)
    return exe


def test_c_match_host_program_builds():
    """no GPU needed: the C test program compiles and links against libkmer_cuda.so and the oracle"""
    assert build_c_test().exists()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,1", "0", "0,0,0"])
def test_multi_gpu_match_c_abi_vs_oracle(devices):
    import torch
    if devices == "0,1" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = build_c_test()
    r = subprocess.run([str(exe), devices], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_sharded_matcher_single_rank_vs_oracle():
    """ShardedMatcher without a process group (world 1): the device-resident match + the hit counters through its buffers."""
    import torch
    import conftest  # noqa: F401
    from kmer_extension_b200 import api, datagen, sharded
    from oracle import oracle as O
    eng = api.KmerCuda(0)
    try:
        sm = sharded.ShardedMatcher(eng)
        m, k = 70001, 12
        col = datagen.synth_kmer_codes(61, m, k)
        consts = datagen.synth_qkmers(62, 9, k, with_n=True) + ["N" * k]
        lo, hi = sm.slice_of(m, 0, 1)
        assert (lo, hi) == (0, m)
        d_codes = torch.from_numpy(col.view(np.int64)).cuda()
        wl = sm.words_per_row(m)
        d_bits = torch.zeros(len(consts) * wl, dtype=torch.int32, device="cuda")
        d_hits = torch.zeros(len(consts), dtype=torch.int64, device="cuda")
        sm.match(api.OP_CONTAINS, d_codes, m, k, consts, d_bits, d_hits)
        full = sm.gather_bits(d_bits, m, m, len(consts)).cpu().numpy()
        got = np.unpackbits(full.view(np.uint32).view(np.uint8), axis=1, bitorder="little")[:, :m].astype(bool)
        hits = d_hits.cpu().numpy()
        for i, p in enumerate(consts):
            want = O.np_match(2, col, k, p).astype(bool)
            assert np.array_equal(got[i], want) and int(hits[i]) == int(want.sum()), p
        with pytest.raises(api.KmerSqlError):
            sm.match(api.OP_CONTAINS, d_codes, m, k, ["acgx"], d_bits, d_hits)
    finally:
        eng.close()


@pytest.mark.gpu
def test_sharded_match_over_nccl_2gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(ROOT / "tests" / "nccl_match_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ))
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "MISMATCH" not in r.stdout

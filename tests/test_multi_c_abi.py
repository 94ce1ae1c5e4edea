"""The multi-GPU entry points of the C ABI driven from a plain C host (tests/c/test_multi.c, no torch / NCCL / Python in the
data path), checked against the C oracle: two contexts on GPU 0 always, two real GPUs when the box has them."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "kmer-extension_b200"


def build_c_test() -> Path:
    from oracle import oracle as O
    O.build(ref=False)
    exe = ROOT / "tests" / "c" / "test_multi"
    cmd = ["gcc", "-O2", "-std=gnu11", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "test_multi.c"), "-o", str(exe),
           f"-L{PKG}", "-lkmer_cuda", f"-L{ROOT / 'oracle'}", "-lkmer_oracle", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{ROOT / 'oracle'}"]
    subprocess.run(cmd, check=True)
    return exe


def test_c_host_program_builds():
    """no GPU needed: the C test program compiles and links against libkmer_cuda.so and the oracle"""
    assert build_c_test().exists()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["0,0", "0,1", "0"])
def test_multi_gpu_c_abi_vs_oracle(devices):
    import torch
    if devices == "0,1" and torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = build_c_test()
    r = subprocess.run([str(exe), devices], capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout)
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0 and "MISMATCH" not in r.stdout, r.stdout[-2000:]

"""Sharded counting over REAL NCCL (2 GPUs, torchrun) against the oracle's table, skewed input included.
Skipped on boxes with fewer than two GPUs (the CPU plumbing test is tests/test_sharded_cpu.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["pull", "nccl"])   # peers' segments read in place over NVLink / NCCL all-to-all
def test_sharded_count_over_nccl_2gpu(exchange):
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731" if exchange == "pull" else "29733", str(ROOT / "tests" / "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, KMER_SHARD_EXCHANGE=exchange))
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0, "sharded count over NCCL disagrees with the oracle (see output)"
    assert "MISMATCH" not in r.stdout

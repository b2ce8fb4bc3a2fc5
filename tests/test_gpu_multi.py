"""Multi-GPU tests (need >= 2 B200 on the box; skipped otherwise): sharded sweep stage + NCCL gather must equal the
single-GPU COO bit for bit, and the row-partitioned distributed LSMR must match the single-GPU solve."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("which", ["taipei", "small"])
def test_sharded_gather_and_distributed_lsmr_match_single_gpu(which):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "scripts", "dist_check.py"),
                        which], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

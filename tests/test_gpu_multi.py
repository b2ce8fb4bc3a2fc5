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


@pytest.mark.parametrize("which,eikonal", [("taipei", "exact"), ("small", "exact"), ("small", "fim")])
def test_sharded_gather_and_distributed_lsmr_match_single_gpu(which, eikonal):
    """both eikonal pipelines: the fast-iterative one relaxes every sweep in a fixed order (one warp per sweep), so its
    sharded result must equal the single-GPU one bit for bit as well"""
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    nproc = 2 if n < 4 else 4
    env = dict(os.environ, DSURF_EIKONAL=eikonal)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
                        "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "scripts", "dist_check.py"),
                        which], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]

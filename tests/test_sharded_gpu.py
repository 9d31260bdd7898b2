"""GPU, N >= 2: the sharded path (one process per GPU over NCCL) equals the single-GPU computation.

Spawns `torch.distributed.run` over every visible GPU with scripts/check_sharded.py: link table / parameters broadcast
from rank 0, each rank computes FK + Jacobian + RNEA + backward on its contiguous row shard, the gathered outputs must be
BIT-identical to rank 0's full-batch computation and the all-reduced link-parameter gradients equal to 1e-4 relative
(fp32 summation order differs).  Skipped on a single-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_equals_single_gpu():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(REPO, "scripts", "check_sharded.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "sharded check ok" in out.stdout

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
    assert "fused peer all-reduce + Adam ok" in out.stdout


def test_fused_adam_single_rank_matches_torch_adam():
    """world = 1: drmb200_allreduce_adam degenerates to a fused Adam step; torch.optim.Adam's arithmetic."""
    from differentiable_robot_model_b200 import parallel
    gen = torch.Generator().manual_seed(0)
    init = torch.randn(257, generator=gen).cuda()
    pa, pb = torch.nn.Parameter(init.clone()), torch.nn.Parameter(init.clone())
    fused = parallel.PeerAllReduceAdam(pa, lr=3e-3, betas=(0.8, 0.95), eps=1e-7)
    ref = torch.optim.Adam([pb], lr=3e-3, betas=(0.8, 0.95), eps=1e-7)
    stream_graph = torch.cuda.CUDAGraph()
    for step in range(5):
        g = torch.randn(257, generator=gen).cuda()
        pa.grad, pb.grad = g.clone(), g.clone()
        fused.step()
        ref.step()
    torch.cuda.synchronize()
    assert float((pa.data - pb.data).abs().max()) < 1e-6 * float(pb.data.abs().max())
    # graph replay advances the device-side step counter like eager launches do
    pa.grad = torch.randn(257, generator=gen).cuda()
    pb.grad = pa.grad.clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(stream_graph, stream=s):
            fused.step()
        before = pa.data.clone()
        stream_graph.replay(); stream_graph.replay()
        s.synchronize()
    ref.step(); ref.step()
    assert not torch.equal(before, pa.data)
    assert float((pa.data - pb.data).abs().max()) < 2e-6 * float(pb.data.abs().max())

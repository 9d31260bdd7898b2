#!/usr/bin/env python
"""torchrun --nproc-per-node N scripts/check_sharded.py : N-GPU correctness of the sharded path.

Rank 0 owns the (learnable) model; the link table is broadcast once over NCCL; every rank computes FK+Jacobian,
RNEA and the backward on its contiguous row shard; the gathered outputs and the all-reduced parameter gradients must
equal rank 0's single-GPU computation on the full batch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import differentiable_robot_model_b200 as drm  # noqa: E402
from differentiable_robot_model_b200 import parallel  # noqa: E402
from differentiable_robot_model_b200.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor  # noqa: E402
from oracle import drm_oracle as O  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234 + rank)                    # different initial parameters per rank on purpose
    m = drm.DifferentiableKUKAiiwa(device=dev)
    m.make_link_param_learnable("iiwa_link_2", "mass", UnconstrainedScalar())
    m.make_link_param_learnable("iiwa_link_4", "com", UnconstrainedTensor(1, 3))
    for p in m.parameters():                          # rank 0's parameters are the truth
        dist.broadcast(p.data, src=0)
    table = parallel.broadcast_link_table(m, src=0)

    B = 100003
    robot = O.load_robot(m.urdf_path, torch.float32)
    q, qd, qdd = (t.to(dev) for t in O.sample_inputs(robot, B, seed=7))     # replicated inputs (same seed)
    G = torch.randn(B, 7, generator=torch.Generator().manual_seed(5)).to(dev)
    lo, hi = parallel.shard_bounds(B, rank, world)

    pos, quat, jl, ja = m.compute_fk_and_jacobian(q[lo:hi], "iiwa_link_ee")
    tau = m.compute_inverse_dynamics(q[lo:hi], qd[lo:hi], qdd[lo:hi])
    (G[lo:hi] * tau).sum().backward()
    n_red = parallel.allreduce_link_param_grads(m)

    def gather(x):
        sizes = [parallel.shard_bounds(B, r, world) for r in range(world)]
        bufs = [torch.empty((h - l,) + tuple(x.shape[1:]), device=dev) for l, h in sizes]
        dist.all_gather(bufs, x.contiguous())
        return torch.cat(bufs)

    full = [gather(t.detach()) for t in (pos, quat, jl, ja, tau)]
    sharded_grads = [p.grad.clone() for p in m.parameters()]
    if rank == 0:
        for p in m.parameters():
            p.grad = None
        ref = list(m.compute_fk_and_jacobian(q, "iiwa_link_ee"))
        tau_ref = m.compute_inverse_dynamics(q, qd, qdd)
        (G * tau_ref).sum().backward()
        for a, b in zip(full, ref + [tau_ref.detach()]):
            assert torch.equal(a, b), "sharded outputs differ from the single-GPU outputs"
        for g_sh, p in zip(sharded_grads, m.parameters()):
            rel = float((g_sh - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-12))
            assert rel < 1e-4, f"all-reduced parameter gradient differs: {rel}"
        print(f"sharded check ok: world={world}, B={B}, {n_red} gradient scalars all-reduced, outputs bit-identical")
    # the fused exchange step (csrc/comm.cu): peer-memory SUM all-reduce + Adam in one kernel, against NCCL all-reduce +
    # torch.optim.Adam on the same per-rank gradients; all ranks must end bit-identical
    P = 91
    gen = torch.Generator().manual_seed(11)
    init = torch.randn(P, generator=gen).to(dev)
    pa = torch.nn.Parameter(init.clone()); pb = torch.nn.Parameter(init.clone())
    fused = parallel.PeerAllReduceAdam(pa, lr=1e-2)
    ref = torch.optim.Adam([pb], lr=1e-2)
    for step in range(6):
        g = torch.randn(P, generator=torch.Generator().manual_seed(1000 * step + rank)).to(dev)
        pa.grad = g.clone()
        fused.step()
        total = g.clone()
        dist.all_reduce(total)
        pb.grad = total
        ref.step()
    torch.cuda.synchronize()
    assert not fused.peer_timeout()
    rel = float((pa.data - pb.data).abs().max() / pb.data.abs().max())
    assert rel < 1e-5, f"fused all-reduce + Adam differs from NCCL + torch Adam: {rel}"
    everyone = [torch.empty_like(pa.data) for _ in range(world)]
    dist.all_gather(everyone, pa.data)
    assert all(torch.equal(e, everyone[0]) for e in everyone), "ranks diverged"
    if rank == 0:
        print(f"fused peer all-reduce + Adam ok: {world} ranks bit-identical, max rel diff vs NCCL + torch Adam {rel:.2e}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

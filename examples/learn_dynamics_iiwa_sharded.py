"""Identify the inertial parameters of all seven Kuka iiwa links from joint torques on a batch that is SHARDED over the
GPUs of one node (BASELINE config 5; the reference has no distributed code -- its loop is
``examples/learn_dynamics_iiwa.py:81-92`` on one device).

    python examples/learn_dynamics_iiwa_sharded.py                                  # one GPU
    torchrun --standalone --nproc-per-node 8 examples/learn_dynamics_iiwa_sharded.py   # eight GPUs, one process each

Every rank owns a contiguous slice of the rows (no data-path collective).  The 21 learnable tensors are fused into ONE flat
``nn.Parameter`` that the link-table kernel reads directly (``model.fuse_learnable_parameters()``), and the only exchange
step -- the SUM of that parameter's gradient over the ranks -- is fused with the Adam update into one kernel over NVLink
peer memory (``parallel.PeerAllReduceAdam``).  The whole iteration is captured once in a CUDA graph and replayed.
"""
import os

import torch
import torch.distributed as dist

from differentiable_robot_model_b200 import DifferentiableKUKAiiwa, DifferentiableRobotModel, parallel
from differentiable_robot_model_b200.data_utils import generate_sine_motion_inverse_dynamics_data
from differentiable_robot_model_b200.rigid_body_params import PositiveScalar, UnconstrainedTensor


def run(n_iters=300, n_data=32768, lr=2e-2, log=print):
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    truth = DifferentiableKUKAiiwa(device=device)
    torch.manual_seed(0)                                            # identical initial guesses on every rank
    student = DifferentiableRobotModel(truth.urdf_path, name="kuka_iiwa", device=device)
    for i in range(1, 8):
        link = f"iiwa_link_{i}"
        student.make_link_param_learnable(link, "mass", PositiveScalar())
        student.make_link_param_learnable(link, "com", UnconstrainedTensor(dim1=1, dim2=3))
        student.make_link_param_learnable(link, "inertia_mat", UnconstrainedTensor(dim1=3, dim2=3))
    flat = student.fuse_learnable_parameters()                      # 7 x (1 + 3 + 9) = 91 values, one tensor
    parallel.broadcast_link_table(student)                          # rank 0's initial values everywhere (no-op on one rank)

    data = generate_sine_motion_inverse_dynamics_data(truth, n_data=n_data, dt=1.0 / 250.0, freq=0.05)
    lo, hi = parallel.shard_bounds(n_data, rank, world)
    q, qd, qdd, tau = (t[lo:hi].contiguous() for t in (data.data["q"], data.data["qd"], data.data["qdd_des"], data.data["tau"]))
    variance = data.var()
    scale = 1.0 / (n_data * tau.shape[1])                           # the loss is a mean over ALL rows: shards add up

    optimiser = parallel.PeerAllReduceAdam(flat, lr=lr)
    flat.grad = torch.zeros_like(flat)
    loss_value = torch.zeros((), device=device)

    def iteration():
        flat.grad.zero_()
        prediction = student.compute_inverse_dynamics(q=q, qd=qd, qdd_des=qdd, include_gravity=True)
        loss = (((prediction - tau) ** 2) / variance).sum() * scale
        loss.backward()
        loss_value.copy_(loss.detach())
        optimiser.step()                                            # SUM over ranks + Adam, one kernel

    stream = torch.cuda.Stream(device=device)
    with torch.cuda.stream(stream):
        for _ in range(3):
            iteration()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            iteration()
        history = []
        for i in range(n_iters):
            graph.replay()
            if i % 50 == 0 or i == n_iters - 1:
                total = loss_value.clone()
                if world > 1:
                    dist.all_reduce(total)                          # logging only
                history.append(float(total))
                if rank == 0:
                    log(f"i: {i} loss: {history[-1]:.6f}")
        stream.synchronize()
    assert not optimiser.peer_timeout()
    if world > 1:
        dist.destroy_process_group()
    return history


if __name__ == "__main__":
    run()

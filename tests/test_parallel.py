"""CPU, world_size 2, gloo: the host-side logic of the N > 1 path (differentiable_robot_model_b200/parallel.py):
row sharding, the single link-table broadcast, and the SUM all-reduce of link-parameter gradients.
No kernel runs here (no GPU); the data path itself has no collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO, urdf_path
from differentiable_robot_model_b200 import parallel


def test_shard_bounds_cover_the_batch_exactly():
    for batch in (0, 1, 7, 8, 65536, 262144, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
                assert hi == lo2 and hi >= lo
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import differentiable_robot_model_b200 as drm
    from differentiable_robot_model_b200.rigid_body_params import UnconstrainedScalar, UnconstrainedTensor

    torch.manual_seed(100 + rank)                       # ranks start with DIFFERENT learnable parameters
    model = drm.DifferentiableRobotModel(urdf_path("iiwa7"), "p")
    model.make_link_param_learnable("iiwa_link_1", "mass", UnconstrainedScalar())
    model.make_link_param_learnable("iiwa_link_2", "trans", UnconstrainedTensor(dim1=1, dim2=3))

    # (1) the single broadcast.  Learnable model: the PARAMETERS are broadcast, so every later (differentiable) table
    # build on every rank starts from rank 0's values
    mine = model._link_table().detach().clone()
    table = parallel.broadcast_link_table(model, src=0)
    gathered = [torch.empty_like(table) for _ in range(world)]
    dist.all_gather(gathered, table)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    assert (rank == 0) == torch.equal(mine, table)
    rebuilt = model._link_table()                       # grad mode on: rebuilt from the (now identical) parameters
    assert rebuilt.requires_grad and torch.equal(rebuilt.detach(), table)
    params = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    all_params = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather(all_params, params)
    assert all(torch.equal(a, all_params[0]) for a in all_params)
    # constant model: the table itself is broadcast and pinned
    const = drm.DifferentiableRobotModel(urdf_path("iiwa7"), "c")
    t = parallel.broadcast_link_table(const, src=0)
    assert const._link_table() is t

    # (2) sharding: shards of a replicated tensor reassemble to the original
    full = torch.arange(11 * 7, dtype=torch.float32).reshape(11, 7)
    part = parallel.shard_rows(full)
    sizes = [parallel.shard_bounds(11, r, world) for r in range(world)]
    assert part.shape[0] == sizes[rank][1] - sizes[rank][0]
    pieces = [torch.empty(hi - lo, 7) for lo, hi in sizes]
    dist.all_gather_object(obj := [None] * world, part)
    assert torch.equal(torch.cat(obj), full)

    # (3) parameter-gradient all-reduce == gradient of the un-sharded batch sum
    for p in model.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    n = parallel.allreduce_link_param_grads(model)
    assert n == 1 + 3
    for p in model.parameters():
        assert torch.equal(p.grad, torch.full_like(p, float(sum(range(1, world + 1)))))
    # a rank whose shard produced no gradient still participates
    for p in model.parameters():
        p.grad = None if rank == 1 else torch.ones_like(p)
    parallel.allreduce_link_param_grads(model)
    for p in model.parameters():
        assert torch.equal(p.grad, torch.full_like(p, float(world - 1)))
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()


def test_world_size_2_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_numa_binding_is_a_safe_no_op_without_nvml():
    """No GPU / NVML in the CPU container: the helper must report False and leave the affinity alone."""
    import os
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_device_numa_node(0) in (False, True)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before

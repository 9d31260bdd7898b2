"""
Multi-GPU plumbing: one process per GPU, the batch of joint configurations is sharded
=======================================================================================
Every row of ``q`` / ``qd`` / ``qdd`` is an independent unit, so the data path needs
NO collective: each rank runs the same kernels on its contiguous slice.  ``torch.distributed`` is
only used for

* ``broadcast_link_table``  -- one broadcast (NCCL on GPUs) of the < 8 KB link table from rank 0 at
  model creation / after an optimiser step on rank 0, so that all ranks compute with bit-identical
  parameters;
* ``allreduce_link_param_grads`` -- one SUM all-reduce of the (tiny) flattened link-parameter
  gradients after ``backward()`` when link parameters are being learned (config 5): parameter
  gradients are batch sums, hence must be summed over the shards.  Gradients w.r.t. the per-row
  inputs stay sharded.

The reference has no distributed code at all; this module is new surface.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced row range [lo, hi) of ``rank``; the first ``batch % world`` ranks get one extra row."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t: torch.Tensor, rank: int = None, world: int = None) -> torch.Tensor:
    """This rank's rows of a replicated ``[B, ...]`` tensor."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def broadcast_link_table(model, src: int = 0) -> torch.Tensor:
    """Make every rank compute with rank ``src``'s link parameters and return the (now identical) link table.

    Constant models: ONE broadcast of the < 8 KB table, pinned as every rank's cached table.  Models with learnable
    link parameters rebuild their table from the parameters on every call, so there the PARAMETERS (and buffers) of the
    parametrisation modules are broadcast instead -- one flat buffer -- and the table follows from them."""
    active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    state = [t for t in list(model.parameters()) + list(model.buffers())]
    if state and model._any_learnable_module():
        flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in state])
        if active:
            dist.broadcast(flat, src=src)
        off = 0
        with torch.no_grad():
            for t in state:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
                off += n
        return model._link_table().detach()
    table = model._link_table().detach().clone()
    if active:
        dist.broadcast(table, src=src)
    model._table_cache = table
    return table


def allreduce_link_param_grads(model) -> int:
    """SUM-all-reduce the gradients of all learnable link parameters in one flat buffer.
    Returns the number of scalars reduced."""
    params = [p for p in model.parameters() if p.requires_grad]
    if not params:
        return 0
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return off


def bind_to_device_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPU cores closest to GPU ``device_index`` (NVML's ideal affinity).

    One process per GPU with host buffers on the box (``drmb200_fk_jacobian_host``): the page-locked buffers should
    live on the NUMA node the GPU's PCIe root hangs off, otherwise every tile crosses the socket interconnect and
    the ranks contend for it.  Call this BEFORE allocating (first-touching) pinned buffers.  Returns False, changing
    nothing, when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False

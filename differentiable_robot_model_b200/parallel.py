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
    model._folded_cache = None
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


class PeerAllReduceAdam:
    """SUM all-reduce of ``param.grad`` over NVLink peer memory FUSED with the Adam update: ONE kernel per step
    (``drmb200_allreduce_adam``, ``csrc/comm.cu``), no NCCL call, CUDA-graph capturable.  For the flat parameter vector of
    ``model.fuse_learnable_parameters()`` on a batch that is sharded over the GPUs of one node: every rank calls
    ``step()`` once per iteration after ``backward()``; all ranks end with bit-identical parameters (rank-ordered sum).
    ``torch.distributed`` is only used once, at construction, to exchange the 64-byte CUDA IPC handles of the inboxes.
    With a single process it degenerates to a fused Adam step.  Same arithmetic as ``torch.optim.Adam`` (defaults)."""

    def __init__(self, param: torch.nn.Parameter, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, group=None):
        import ctypes
        from . import engine
        if not param.is_cuda or param.dtype != torch.float32 or not param.is_contiguous():
            raise RuntimeError("PeerAllReduceAdam needs a contiguous fp32 CUDA parameter")
        self.param, self.lr, self.betas, self.eps = param, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(param.data)
        self.exp_avg_sq = torch.zeros_like(param.data)
        active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if active else 1
        self.rank = dist.get_rank(group) if active else 0
        self._lib = engine.lib()
        self._comm = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(param.device):
            engine._check(self._lib.drmb200_comm_create(self.rank, self.world, max(1, param.numel()), ctypes.byref(self._comm),
                                                        handle), "drmb200_comm_create")
            if self.world > 1:
                mine = torch.tensor(list(handle), dtype=torch.uint8, device=param.device)
                everyone = [torch.empty_like(mine) for _ in range(self.world)]
                dist.all_gather(everyone, mine, group=group)
                blob = bytes(torch.cat(everyone).cpu().tolist())
                engine._check(self._lib.drmb200_comm_connect(self._comm, blob), "drmb200_comm_connect")
                dist.barrier(group=group)                 # every inbox is mapped before anyone writes into it

    def zero_grad(self, set_to_none: bool = False):
        if self.param.grad is not None:
            if set_to_none:
                self.param.grad = None
            else:
                self.param.grad.zero_()

    def step(self):
        from . import engine
        g = self.param.grad
        if g is None:
            raise RuntimeError("PeerAllReduceAdam.step(): the parameter has no gradient (every rank must step every iteration)")
        g = g.contiguous()
        with torch.cuda.device(self.param.device):
            rc = self._lib.drmb200_allreduce_adam(self._comm, engine._ptr(self.param.data), engine._ptr(g), engine._ptr(self.exp_avg),
                                                  engine._ptr(self.exp_avg_sq), self.param.numel(), self.lr, self.betas[0],
                                                  self.betas[1], self.eps, engine._stream())
        engine._check(rc, "drmb200_allreduce_adam")

    def peer_timeout(self) -> bool:
        """True if a peer failed to show up within ~2 s in some step (synchronises the device)."""
        return self._lib.drmb200_comm_error(self._comm) != 0

    def __del__(self):
        try:
            if getattr(self, "_comm", None) is not None and self._comm.value:
                self._lib.drmb200_comm_destroy(self._comm)
                self._comm = None
        except Exception:
            pass


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

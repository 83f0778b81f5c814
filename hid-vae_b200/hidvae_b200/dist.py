"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch; gloo in the CPU
tests).  Replaces what the reference gets implicitly from HuggingFace Accelerate -> DDP (train_hidvae.py:186-189,
630-632, 709): parameters are broadcast once, and every step ends in ONE all-reduce of ONE flat gradient buffer
(29 MB for the Amazon model: latency-bound on NVSwitch, so one launch beats DDP's bucketed many)."""
import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank); initialises the default process group when launched by torchrun."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class FlatGradAllReduce:
    """Gives every parameter a `.grad` that is a view into one flat fp32 buffer, so that autograd accumulates
    straight into it and the data-parallel exchange is a single collective with no packing copies."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def check_views(self) -> None:
        """autograd / optimizers that set grads to None would silently detach a parameter from the flat buffer."""
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + self.flat.numel() * self.flat.element_size()
        for p in self.params:
            if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer "
                                   "(use optimizer.zero_grad(set_to_none=False) or FlatGradAllReduce.zero())")

    def all_reduce(self, async_op: bool = False):
        if world_size(self.group) == 1:
            return None
        work = dist.all_reduce(self.flat, group=self.group, async_op=async_op)
        if self.average and not async_op:
            self.flat.div_(world_size(self.group))
        return work

    def finish(self, work) -> None:
        if work is not None:
            work.wait()
            if self.average:
                self.flat.div_(world_size(self.group))


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> None:
    if world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        if t is not None:
            dist.broadcast(t.data, src=src, group=group)


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous [lo, hi) of `n` items owned by `rank` (bulk semantic-ID assignment shards by items)."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)

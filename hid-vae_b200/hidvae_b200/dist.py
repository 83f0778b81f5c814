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
    straight into it and the data-parallel exchange is a single collective with no packing copies.
    Every parameter passed in therefore has a (possibly zero) gradient on every step: freeze parameters that cannot
    receive one (`requires_grad_(False)`) so that they stay out of the buffer and the optimizer leaves them alone."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, average: bool = True) -> None:
        self.params = [p for p in params if p.requires_grad]
        self.group, self.average = group, average
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off: off + p.numel()].view_as(p))
            p.grad = self.views[-1]
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def backward(self, loss: torch.Tensor) -> None:
        """`loss.backward()` with the gradients ADDED to the flat buffer by one multi-tensor kernel.  A plain backward() on
        parameters whose `.grad` is a view costs one `add_` launch per parameter (autograd accumulates into a defined grad:
        176 launches of 1.7 us in a HiD-VAE step, 9 % of the step at batch 128); here the grads are detached first, autograd
        hands over its freshly computed tensors, and `_foreach_add_` folds them into the views, which are then bound again
        (optimizer, all_reduce and check_views see the same views as before; parameters without a gradient keep zeros)."""
        for p in self.params:
            p.grad = None
        loss.backward()
        dst, src = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                dst.append(v)
                src.append(p.grad)
            p.grad = v
        if dst:
            torch._foreach_add_(dst, src)

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


def broadcast_buffers(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Rank `src`'s buffers (BatchNorm running statistics) to every rank -- DDP does this before every forward; here it
    runs before evaluation and checkpointing, the two places where the buffers are read."""
    if world_size(group) == 1:
        return
    for t in module.buffers():
        if t is not None and t.is_floating_point():
            dist.broadcast(t.data, src=src, group=group)


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous [lo, hi) of `n` items owned by `rank` (bulk semantic-ID assignment shards by items)."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


class PeerAllReduce:
    """In-place sum of a small fp32 tensor over all ranks as ONE kernel over NVLink / NVSwitch peer memory
    (`hv_peer_allreduce`, csrc/peer_allreduce.cu): every rank pushes its contribution into an inbox on every peer and sums
    the inboxes in rank order, so all ranks end with bit-identical results.  For the latency-bound exchanges of the
    quantiser (98 KB of codebook gradient per step); large buffers belong to NCCL (`FlatGradAllReduce`).
    The inboxes and flags are a symmetric allocation (torch.distributed._symmetric_memory: CUDA peer mappings between the
    processes of one node).  No fallback: a node without peer access fails in the constructor.

    Ordering: consecutive calls on one instance must not overlap (the two-parity inbox reuse relies on it).  Calls made on
    different streams are therefore chained with an event; inside a CUDA-graph capture the caller keeps them on one
    stream.  A rank that never arrives does not trap the kernel: after `timeout_ms` the call completes with an invalid
    result and `check()` (a host synchronisation) raises, naming the missing rank."""

    def __init__(self, numel: int, device, group=None) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        from hidvae_b200._lib import lib
        if numel % 4:
            raise ValueError("PeerAllReduce: numel must be a multiple of 4 (rows are moved as 16-byte vectors)")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.n = dist.get_world_size(self.group), dist.get_rank(self.group), numel
        n_chunks = int(lib.hv_peer_allreduce_chunks(numel))
        name = self.group.group_name
        self.inbox = symm_mem.empty(2 * self.world * numel, dtype=torch.float32, device=device)
        self.flags = symm_mem.empty(2 * self.world * n_chunks, dtype=torch.int32, device=device)
        self.flags.zero_()
        h_inbox, h_flags = symm_mem.rendezvous(self.inbox, name), symm_mem.rendezvous(self.flags, name)
        self.inbox_ptrs = torch.tensor([int(p) for p in h_inbox.buffer_ptrs], dtype=torch.int64, device=device)
        self.flag_ptrs = torch.tensor([int(p) for p in h_flags.buffer_ptrs], dtype=torch.int64, device=device)
        self.n_chunks = n_chunks
        self.seq = torch.zeros(n_chunks + 1, dtype=torch.int32, device=device)   # call counters + sticky status word
        self._last = None
        torch.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's flags are zero before anybody pushes

    def __call__(self, t: torch.Tensor) -> torch.Tensor:
        from hidvae_b200._lib import check, lib
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.n:
            raise ValueError(f"PeerAllReduce: expected a contiguous fp32 tensor of {self.n} elements")
        with torch.cuda.device(t.device):
            stream = torch.cuda.current_stream(t.device)
            capturing = torch.cuda.is_current_stream_capturing()
            if self._last is not None and not capturing:
                stream.wait_event(self._last)      # the previous call (possibly on another stream) has left the inboxes
            check(lib.hv_peer_allreduce(t.data_ptr(), t.data_ptr(), self.n, self.inbox_ptrs.data_ptr(), self.flag_ptrs.data_ptr(),
                                        self.seq.data_ptr(), self.rank, self.world, stream.cuda_stream))
            if not capturing:
                self._last = torch.cuda.Event()
                self._last.record(stream)
        return t

    @staticmethod
    def set_timeout_ms(ms: int) -> None:
        from hidvae_b200._lib import lib
        lib.hv_peer_allreduce_set_timeout_ms(int(ms))

    def check(self) -> None:
        """Host synchronisation: raises HidvaeError if any call so far gave up waiting for a rank."""
        import ctypes

        from hidvae_b200._lib import check, lib
        word = int(self.seq[self.n_chunks].item()) & 0xFFFFFFFF
        check(lib.hv_peer_allreduce_status(ctypes.c_uint32(word), None, None))

"""K-means codebook init with the interface of the reference's init/kmeans.py (`kmeans_init_`, `Kmeans`,
`KmeansOutput`) over the CUDA kernels of libhidvae_b200.so.

Same algorithm and stop rule as the reference (init/kmeans.py:34-77): distinct random rows as initial
centroids drawn with `np.random.choice` from NumPy's global RNG, assignment by the exact difference-form
distance, per-cluster means, empty clusters re-seeded from `torch.randint` rows in cluster order, loop until
max ||c_new - c_old||_2 < stop_threshold.  What differs is where the work happens: one assignment kernel (no
[N, K, D] temporary), one deterministic segmented-sum kernel (no K-iteration Python loop with K device syncs)
and one finalize kernel per Lloyd iteration, and ONE two-float read-back per iteration for the stop test.

Data-parallel use (`process_group=`): every rank passes its own shard of the rows; sums and counts are
all-reduced each iteration (33 KB at K=256, D=32), initial rows and re-seeds are drawn on rank 0 and broadcast, so
all ranks end with identical centroids -- unlike the reference, whose ranks each run an unsynchronised k-means
(SURVEY.md section 8e).
"""
from typing import NamedTuple, Optional

import numpy as np
import torch
from torch import Tensor

from hidvae_b200 import ops


class KmeansOutput(NamedTuple):
    centroids: Tensor
    assignment: Tensor


def kmeans_init_(tensor: Tensor, x: Tensor, process_group=None) -> None:
    assert tensor.dim() == 2
    assert x.dim() == 2
    with torch.no_grad():
        out = Kmeans(k=tensor.shape[0], process_group=process_group).run(x)
        tensor.data.copy_(out.centroids)


class Kmeans:
    def __init__(self, k: int, max_iters: Optional[int] = None, stop_threshold: float = 1e-10, process_group=None) -> None:
        self.k = k
        self.iters = max_iters
        self.stop_threshold = stop_threshold
        self.process_group = process_group
        self.centroids = None
        self.assignment = None
        self.n_iters = 0

    # ---- distributed plumbing (no-ops for a single process) ---------------------------------------------------
    def _dist(self):
        if self.process_group is None:
            return None
        import torch.distributed as dist
        return dist if dist.get_world_size(self.process_group) > 1 else None

    def _row_offsets(self, n_local: int, device):
        """(first global row of this rank, total rows)."""
        dist = self._dist()
        if dist is None:
            return 0, n_local
        sizes = torch.zeros(dist.get_world_size(self.process_group), dtype=torch.int64, device=device)
        sizes[dist.get_rank(self.process_group)] = n_local
        dist.all_reduce(sizes, group=self.process_group)
        sizes = sizes.tolist()
        return sum(sizes[: dist.get_rank(self.process_group)]), sum(sizes)

    def _fetch_rows(self, x: Tensor, global_idx: Tensor, first: int) -> Tensor:
        """Rows of the GLOBAL matrix by global index: every rank contributes the rows it owns."""
        dist = self._dist()
        if dist is None:
            return x[global_idx.to(x.device)]
        idx = global_idx.to(x.device)
        local = idx - first
        mine = (local >= 0) & (local < x.shape[0])
        rows = torch.zeros((idx.numel(), x.shape[1]), dtype=x.dtype, device=x.device)
        rows[mine] = x[local[mine]]
        dist.all_reduce(rows, group=self.process_group)
        return rows

    def _broadcast_idx(self, idx: Tensor, device) -> Tensor:
        dist = self._dist()
        if dist is None:
            return idx
        idx = idx.to(device)
        dist.broadcast(idx, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
        return idx

    # ---- reference-shaped steps ---------------------------------------------------------------------------------
    def _init_centroids(self, x: Tensor) -> None:
        first, total = self._row_offsets(x.shape[0], x.device)
        self._first, self._total = first, total
        # np.random.choice raises ValueError when total < k, like the reference (init/kmeans.py:38)
        init_idx = torch.from_numpy(np.asarray(np.random.choice(total, self.k, replace=False), dtype=np.int64))
        init_idx = self._broadcast_idx(init_idx, x.device)
        self.centroids = self._fetch_rows(x, init_idx, first).contiguous().clone()
        self.assignment = None

    def _update_centroids(self, x: Tensor) -> float:
        """One Lloyd iteration; returns max_c ||c_new - c_old||_2 (the quantity of init/kmeans.py:68)."""
        dist = self._dist()
        # Assignment = the fused distance + argmin kernel with one level.  Small problems use its exact-fp32 CUDA-core
        # variant in the reference's difference form sum (x - c)^2 (init/kmeans.py:44-47: assignments then match the
        # reference bit for bit); from K * N > 2^24 on (K = 4096 at 65,536 rows is 17 G multiply-adds) the tcgen05 variant
        # takes over -- GEMM form, fp32-grade scores, same near-tie policy as the quantiser (a row may take the other of
        # two centroids whose distances differ by < 1e-5 relative).  Both are deterministic functions of (x, centroids).
        exact = self.k * x.shape[0] <= (1 << 24) or not ops.workspace_bytes(x.shape[1], self.k, 1)
        new_assign = ops.kmeans_assign(x, self.centroids, exact_diff_form=exact)
        sums, counts, _changed = ops.kmeans_accumulate(x, new_assign, self.k, self.assignment)
        if dist is not None:
            dist.all_reduce(sums, group=self.process_group)
            dist.all_reduce(counts, group=self.process_group)
        old = self.centroids.clone()
        stats = ops.kmeans_finalize(sums, counts, self.centroids, reseed_rows=None)
        shift, n_empty = stats.tolist()  # the one host read-back of the iteration
        if n_empty > 0:
            if self._total == 0:
                raise ValueError("Can not choose random element from x, x is empty")
            empty = torch.nonzero(counts == 0).view(-1)
            # one torch.randint per empty cluster, in cluster order (init/kmeans.py:54-56)
            draws = torch.cat([torch.randint(0, self._total, (1,)) for _ in range(empty.numel())])
            draws = self._broadcast_idx(draws, x.device)
            self.centroids[empty] = self._fetch_rows(x, draws, self._first)
            shift = float(torch.norm(self.centroids - old, dim=1).max())
        self.assignment = new_assign
        return shift

    def run(self, x: Tensor) -> KmeansOutput:
        if not x.is_cuda:
            raise RuntimeError("Kmeans.run: x must be a CUDA tensor (hidvae_b200 has no CPU fallback)")
        x = x.detach().float().contiguous()
        self._init_centroids(x)
        i = 0
        self.n_iters = 0
        while self.iters is None or i < self.iters:
            shift = self._update_centroids(x)
            self.n_iters += 1
            if shift < self.stop_threshold:
                break
            i += 1
        return KmeansOutput(centroids=self.centroids, assignment=self.assignment)

"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of init/kmeans.py (Lloyd's algorithm).

Follows init/kmeans.py:34-77: random distinct rows as initial centroids (`np.random.choice`, :38), assignment
by argmin of the exact sum((x - c)^2) table (:44-47), per-cluster mean in cluster order with an empty cluster
reseeded from a `torch.randint` row (:52-60), loop until max ||c_new - c_old||_2 < stop_threshold (:63-72).
The two random sources can be injected so the CUDA path can be driven with identical draws.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch
from torch import Tensor


class KmeansResult(NamedTuple):
    centroids: Tensor   # [K, D]
    assignment: Tensor  # [N] int64
    n_iters: int        # number of centroid updates performed


def assign_exact(x: Tensor, centroids: Tensor, chunk: int = 4096) -> Tensor:
    """argmin_k sum_d (x - c_k)^2, the difference form of init/kmeans.py:44-47 (chunked over rows only to
    bound the [n, K, D] temporary; each row's arithmetic is unchanged)."""
    out = []
    for s in range(0, x.shape[0], chunk):
        diff = x[s:s + chunk].unsqueeze(1) - centroids.unsqueeze(0)
        out.append((diff ** 2).sum(axis=2).min(axis=1).indices)
    return torch.cat(out) if out else torch.empty(0, dtype=torch.int64)


def lloyd_update(x: Tensor, centroids: Tensor, draw_row: Callable[[], int]):
    """One `_update_centroids` (init/kmeans.py:43-61).  Returns (new_centroids, assignment, n_empty)."""
    k = centroids.shape[0]
    assignment = assign_exact(x, centroids)
    new_c = centroids.clone()
    n_empty = 0
    for c in range(k):
        members = assignment == c
        if not members.any():
            if x.size(0) == 0:
                raise ValueError("Can not choose random element from x, x is empty")
            new_c[c] = x[draw_row()]
            n_empty += 1
        else:
            new_c[c] = x[members].mean(axis=0)
    return new_c, assignment, n_empty


def kmeans_run(
    x: Tensor,
    k: int,
    max_iters: Optional[int] = None,
    stop_threshold: float = 1e-10,
    init_idx: Optional[np.ndarray] = None,
    draw_row: Optional[Callable[[], int]] = None,
) -> KmeansResult:
    """`Kmeans(k, max_iters, stop_threshold).run(x)` (init/kmeans.py:63-77)."""
    n = x.shape[0]
    if init_idx is None:
        init_idx = np.random.choice(n, k, replace=False)  # raises ValueError if n < k, like the reference
    if draw_row is None:
        draw_row = lambda: int(torch.randint(0, n, (1,)))
    centroids = x[torch.as_tensor(init_idx, dtype=torch.int64), :].clone()
    assignment = None
    i = 0
    updates = 0
    while max_iters is None or i < max_iters:
        old = centroids
        centroids, assignment, _ = lloyd_update(x, centroids, draw_row)
        updates += 1
        if torch.norm(centroids - old, dim=1).max() < stop_threshold:
            break
        i += 1
    return KmeansResult(centroids, assignment, updates)


def inertia(x: Tensor, centroids: Tensor) -> float:
    """sum of squared distances to the nearest centroid (used to compare runs whose RNG differs)."""
    total = 0.0
    for s in range(0, x.shape[0], 4096):
        diff = x[s:s + 4096].unsqueeze(1).double() - centroids.unsqueeze(0).double()
        total += float((diff ** 2).sum(axis=2).min(axis=1).values.sum())
    return total

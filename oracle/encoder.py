"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's encoder MLP (the step in front of the quantiser).

Follows /root/reference/modules/encoder.py:23-36 (bias-free nn.Linear + nn.SiLU stack, L2NormalizationLayer tail when
`normalize`) and /root/reference/modules/normalize.py:7-8 (F.normalize, eps 1e-12), with the same ATen CPU operators in
the same order.  Pinned by tests/golden/encoder.npz, recorded from the reference's own `MLP` (oracle/make_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor


def seeded_weights(dims: Sequence[int], seed: int) -> List[Tensor]:
    """Linear weights [out, in] with nn.Linear's default bound 1/sqrt(in), drawn from a seeded CPU generator (the
    fixture stores the seed instead of 2.2 MB of weights)."""
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(o, i, generator=g) * 2.0 - 1.0) / (i ** 0.5) for i, o in zip(dims[:-1], dims[1:])]


def mlp_forward(x: Tensor, weights: Sequence[Tensor], normalize: bool) -> Tensor:
    """encoder.py:24-32,36: Linear(bias=False), SiLU between layers (none after the last), optional L2 norm."""
    h = x
    for l, w in enumerate(weights):
        h = F.linear(h, w)
        if l != len(weights) - 1:
            h = F.silu(h)
    return F.normalize(h, p=2, dim=-1, eps=1e-12) if normalize else h

"""Row L2 normalisation used as the codebook `out_proj` tail and the encoder/decoder tail.
Interface of the reference's modules/normalize.py:7-19 (`l2norm`, `L2NormalizationLayer`)."""
import torch.nn.functional as F
from torch import Tensor, nn


def l2norm(x: Tensor, dim: int = -1, eps: float = 1e-12) -> Tensor:
    return F.normalize(x, p=2, dim=dim, eps=eps)


class L2NormalizationLayer(nn.Module):
    def __init__(self, dim: int = -1, eps: float = 1e-12) -> None:
        super().__init__()
        self.dim, self.eps = dim, eps

    def forward(self, x: Tensor) -> Tensor:
        return l2norm(x, self.dim, self.eps)

    def extra_repr(self) -> str:
        return f"dim={self.dim}, eps={self.eps}"

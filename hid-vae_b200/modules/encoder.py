"""Bias-free Linear + SiLU stack with an optional L2-norm tail (reference modules/encoder.py:7-36).

Not part of the hot path: plain cuBLAS GEMMs through PyTorch.  State-dict keys (`mlp.<i>.weight`) match the
reference so its checkpoints load."""
from typing import List

from torch import Tensor, nn

from modules.normalize import L2NormalizationLayer


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dims: List[int], out_dim: int, dropout: float = 0.0,
                 normalize: bool = False) -> None:
        super().__init__()
        self.input_dim, self.hidden_dims, self.out_dim, self.dropout = input_dim, hidden_dims, out_dim, dropout
        widths = [input_dim, *hidden_dims, out_dim]
        stack = []
        for pos in range(len(widths) - 1):
            stack.append(nn.Linear(widths[pos], widths[pos + 1], bias=False))
            if pos < len(widths) - 2:
                stack.append(nn.SiLU())
                if dropout != 0:
                    stack.append(nn.Dropout(dropout))
        stack.append(L2NormalizationLayer() if normalize else nn.Identity())
        self.mlp = nn.Sequential(*stack)

    def forward(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == self.input_dim, f"Invalid input dim: Expected {self.input_dim}, found {x.shape[-1]}"
        return self.mlp(x)

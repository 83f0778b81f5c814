"""Bias-free Linear + SiLU stack with an optional L2-norm tail (reference modules/encoder.py:7-36).

State-dict keys (`mlp.<i>.weight`) match the reference so its checkpoints load.  Training (autograd) runs the PyTorch
layers.  Inference passes on CUDA -- `HRqVae.encode` under `torch.no_grad()`: eval encode, `precompute_corpus_ids`,
`predict_tags` -- run ONE fused tcgen05 kernel (`hv_encoder_forward`, csrc/enc_mlp.cu) when the widths have an
instantiation (the gin shape 768-512-256-128-32): fp16 operands (TF32-grade significand, the precision of the
reference's own GPU path, modules/h_rqvae.py:21), fp32 accumulation and activations, hidden activations never leave the
SM.  `inference_precision` selects the numerics of those passes:
    "fused"  the kernel above (default)
    "tf32"   PyTorch layers with TF32 matmuls (what the reference runs on a GPU)
    "fp32"   PyTorch layers with fp32 matmuls (bit-comparable with the CPU reference)
"""
import contextlib
from typing import List

import torch
from torch import Tensor, nn

from modules.normalize import L2NormalizationLayer

INFERENCE_PRECISIONS = ("fused", "tf32", "fp32")


@contextlib.contextmanager
def _matmul_tf32(enabled: bool):
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = bool(enabled)
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


class MLP(nn.Module):
    def __init__(self, input_dim: int, hidden_dims: List[int], out_dim: int, dropout: float = 0.0,
                 normalize: bool = False) -> None:
        super().__init__()
        self.input_dim, self.hidden_dims, self.out_dim, self.dropout = input_dim, hidden_dims, out_dim, dropout
        self.normalize = normalize
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
        self.inference_precision = "fused"
        self.precise_silu = False
        self._image, self._image_key = None, None

    def _linear_weights(self) -> List[Tensor]:
        return [m.weight for m in self.mlp if isinstance(m, nn.Linear)]

    def fused_available(self, x: Tensor) -> bool:
        """The fused kernel serves this call: CUDA fp32 (or fp16) rows, no autograd, no active dropout, widths instantiated."""
        if not x.is_cuda or x.dtype not in (torch.float32, torch.float16) or x.dim() < 2 or torch.is_grad_enabled():
            return False
        if self.training and self.dropout != 0:
            return False
        from hidvae_b200 import ops
        return ops.encoder_supported([self.input_dim, *self.hidden_dims, self.out_dim])

    def _weight_image(self):
        """fp16 tensor-core image of the current weights, re-packed whenever a weight changed (version counters)."""
        from hidvae_b200 import ops
        ws = self._linear_weights()
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if self._image is None or key != self._image_key:
            self._image, self._image_key = ops.encoder_pack(ws), key
        return self._image

    def forward(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == self.input_dim, f"Invalid input dim: Expected {self.input_dim}, found {x.shape[-1]}"
        if self.inference_precision not in INFERENCE_PRECISIONS:
            raise ValueError(f"inference_precision must be one of {INFERENCE_PRECISIONS}, got {self.inference_precision!r}")
        if self.inference_precision == "fused" and self.fused_available(x):
            from hidvae_b200 import ops
            z = ops.encoder_forward(x.reshape(-1, self.input_dim), self._weight_image(), normalize=self.normalize,
                                    precise_silu=self.precise_silu)
            return z.reshape(*x.shape[:-1], self.out_dim)
        x = x.float()   # (fp16 items are served by the fused kernel only; every other path computes from fp32)
        if x.is_cuda and not torch.is_grad_enabled() and self.inference_precision != "fused":
            with _matmul_tf32(self.inference_precision == "tf32"):
                return self.mlp(x)
        return self.mlp(x)
